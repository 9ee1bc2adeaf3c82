// K1, fourth generation (bf16 tensor-core mode): Conv2d(obs->16, 7x7, stride 3) + bias + ReLU + MaxPool(3) with the
// operand roles SWAPPED and the pooling done in registers. Replaces cnn_base[0:3] of
// /root/reference/src/architectures/nets.py:18-20.
//
// The earlier generations (conv1_tc.cu) put the pixels on the M side: an accumulator row (= TMEM lane = thread) was one
// (conv row, 4-column group) and held 64 (j, co) columns, so BOTH pooling directions crossed threads and every
// accumulator went through a 34 KB shared-memory staging buffer -- the kernel was bound by that epilogue (22 of 28 us
// with the MMAs switched off, profiles/README.md). Here
//     D[(s, co, j), (oy, g)] = sum_{ci, ky, p<16}  Wt[(co,j)][(ci,ky,p)] * in[s + ci][3*oy+ky][12*g + p]
//   * A operand (M = 128) = the Toeplitz WEIGHTS of two consecutive samples of the sliding window: the plane that is
//     channel ci of sample s is channel ci-1 of sample s+1, so with the weight blocks ordered by DESCENDING channel
//     (zero blocks at both ends) the 128 rows [W(ci) ; W(ci-1)] are one contiguous slice of the image;
//   * B operand (N = 128) = the 126 pixels (6 conv rows x 21 groups) of a tile: the same strided views of the
//     Toeplitz-ready planes (stage.cu) the earlier generations used as A;
//   * an accumulator LANE is one (sample, co, j) and holds all 126 pixels of the tile in its columns, so the vertical
//     3-max is register-local, and the horizontal 3-max is a quad exchange (the four j of a channel sit in adjacent
//     lanes: 12 conv columns = 3 groups x 4 j = 4 pooling windows, lane j takes window j). No shared-memory staging,
//     no block barrier: every epilogue warp runs on its own.
// M=128 x N=128 x K=16 costs 64 cycles = the tensor-core rate (tools/mma_bench.py); the pair blocks waste 1 slot in 5.
// Planes stream along the sample axis for a whole run of samples of one tile row (S + 3 loads for S samples).
// All mbarrier waits are bounded and raise a device flag instead of hanging.
#include <stdlib.h>
#include "bc_common.cuh"
#include "tc05.cuh"
#include "pack.cuh"
#include "trace.cuh"

namespace c1f4 {


constexpr int NG = 21;                       // groups of 4 output columns per conv row
constexpr int TILES_PER_FRAME = 14;          // tile = 6 conv rows x 84 columns of one frame (= 2 pooled rows)
constexpr int NISS = 3;                      // MMA issuer warps 1..3; pair pc is issued by warp 1 + pc % NISS (a queued tcgen05.mma pins its uniform registers:
                                             // the next visit's descriptor set-up of the SAME warp waits for them, so short visits need several issuers)
constexpr int NEG = 3;                       // epilogue groups of 4 warps (one warp per TMEM lane quadrant); pair pc goes to group pc % NEG
constexpr int NTHREADS = (4 + 4 * NEG) * 32;  // warp 0 loader + TMEM alloc, 1-3 MMA issuers, 4.. epilogue (128 registers per thread)
constexpr int ROWB = 336;                    // 21 groups x 16 B
constexpr int PIECE0 = 8 * ROWB, PIECE12 = 7 * ROWB;      // class 0 feeds ky 0,3,6 (8 rows), classes 1,2 feed two ky (7 rows)
constexpr int SLOT_BYTES = 2 * PIECE0 + 4 * PIECE12;      // 14784
constexpr int TP_PIECE_BYTES = 86 * ROWB;                 // piece stride inside a TP plane in HBM
__host__ __device__ constexpr int piece_off(int c, int h) { return c == 0 ? h * PIECE0 : 2 * PIECE0 + (c - 1) * 2 * PIECE12 + h * PIECE12; }
__host__ __device__ constexpr int px_off(int ky) { return piece_off(ky % 3, 0) + (ky / 3) * ROWB; }
constexpr int NSLOT = 8;
constexpr int W_BYTES = ctc::kC1V4Bytes;     // 36 blocks of 64 rows x 16 k
constexpr int OFF_W = 0;
constexpr int OFF_RING = OFF_W + W_BYTES;
constexpr int OFF_P = (OFF_RING + NSLOT * SLOT_BYTES + 64 + 127) / 128 * 128;   // 64 B: the over-read of the last slice
constexpr int PW_BYTES = 56 * 16 + 56 * 8;   // per epilogue warp: one (sample, 8 channels) tile as bf16 P8 + its routing
constexpr int OFF_BAR = OFF_P + 4 * NEG * PW_BYTES;
constexpr int NBAR = 1 + 2 * NSLOT + 4 + 4;
constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
constexpr int TMEM_COLS = 512;               // 4 accumulators of 128 columns
static_assert(SMEM_BYTES <= 227 * 1024, "conv1 forward (gen 4) shared memory");
static_assert(OFF_BAR % 8 == 0 && PW_BYTES % 16 == 0, "alignment");

// contiguous balanced range of the (ty, b) sample-tile list, cut into runs of consecutive samples of one tile row
struct RunIter {
    int i, hi, B; bool sliding;
    __device__ RunIter(int B_, bool sliding_) : B(B_), sliding(sliding_) {
        const long long T = (long long)B_ * TILES_PER_FRAME;
        i = (int)(T * blockIdx.x / gridDim.x);
        hi = (int)(T * (blockIdx.x + 1) / gridDim.x);
    }
    __device__ bool next(int& ty, int& b0, int& S) {
        if (i >= hi) return false;
        ty = i / B; b0 = i - ty * B;
        S = sliding ? min(B - b0, hi - i) : 1;       // a materialised batch shares no planes: one sample per run
        i += S;
        return true;
    }
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv1_fwd4_kernel(const __nv_bfloat16* __restrict__ x, int64_t sn, int64_t sc, const uint8_t* __restrict__ wimg,
                  const float* __restrict__ bias, float* __restrict__ y, uint8_t* __restrict__ amax,
                  __nv_bfloat16* __restrict__ ybf, uint8_t* __restrict__ amax_p8, int B, int* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* w_full = bars;
    uint64_t* slot_full = bars + 1;                  // [NSLOT]
    uint64_t* slot_empty = bars + 1 + NSLOT;         // [NSLOT] all issuers
    uint64_t* t_full = bars + 1 + 2 * NSLOT;         // [4] accumulator ring
    uint64_t* t_empty = bars + 5 + 2 * NSLOT;        // [4] the 4 warps that drain an accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool sliding = sn == sc;                   // plane sharing needs the sliding window

    if (threadIdx.x == 0) {
        tc05::mbar_init(w_full, 1);
        for (int i = 0; i < NSLOT; ++i) { tc05::mbar_init(slot_full + i, 1); tc05::mbar_init(slot_empty + i, NISS); }
        for (int i = 0; i < 4; ++i) { tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
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
    // the weight operand image is written only by kernels that release their dependents after their last write (Adam / pack:
    // abi.cu, conv_tc.cu), so it is fetched here, under the previous kernel's tail, and not after the wait
    if (threadIdx.x == 0) {
        tc05::mbar_expect_tx(w_full, W_BYTES);
        tc05::bulk_g2s(smem + OFF_W, wimg, W_BYTES, w_full);
    }
    tc05::pdl_trigger();
    tc05::pdl_wait();                    // everything above overlapped the previous kernel's tail; global memory from here on

    if (warp == 0) {
        // ------------------------------------------------------------------ loader: 6 bulk copies per plane, one per lane
        RunIter it(B, sliding);
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
                if (lane == 0) tc05::mbar_expect_tx(slot_full + slot, SLOT_BYTES);
                TRACE(1, k);
                __syncwarp();
                if (lane < 6)
                    tc05::bulk_g2s(smem + OFF_RING + slot * SLOT_BYTES + dst_off, src0 + (int64_t)j * sc * 2 + src_off, nbytes, slot_full + slot);
            }
        }
    } else if (warp <= NISS) {
        // ------------------------------------------------------------------ MMA issuers: issuer p owns the pairs with pair counter % NISS == p
        // Pair q of a run = samples 2q, 2q+1; plane j of the run is channel d = j - 2q of the first and d - 1 of the second,
        // so it meets the pair for d = 0..4 with A = weight blocks [ci = d ; ci = d - 1] (ci = 4 and ci = -1 are zero blocks).
        // Every issuer observes every plane in ring order and every issuer releases it (slot_empty counts NISS): an issuer
        // that waited only for the planes of its own pairs could be a whole ring cycle ahead of the loader and take the
        // parity of an OLDER use of the slot for its plane (measured: the materialised-batch path failed exactly so).
        const uint32_t par = (uint32_t)(warp - 1);
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, 128, 0, 0);
        const uint32_t ring = tc05::smem_u32(smem + OFF_RING);
        // descriptors as (low word = start address >> 4 and LBO, high word = SBO / version): only the low word moves
        const uint64_t pd_c0 = tc05::smem_desc(ring, PIECE0, 128, tc05::SW_NONE);     // pixels: LBO = distance between the K halves
        const uint64_t pd_c12 = tc05::smem_desc(ring, PIECE12, 128, tc05::SW_NONE);
        const uint64_t wd0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_W), 128, 256, tc05::SW_NONE);
        const uint32_t p0_lo = (uint32_t)pd_c0, p12_lo = (uint32_t)pd_c12, p_hi = (uint32_t)(pd_c0 >> 32);
        const uint32_t w_lo = (uint32_t)wd0, w_hi = (uint32_t)(wd0 >> 32);
        bool ok = tc05::mbar_wait(w_full, 0, err);
        RunIter it(B, sliding);
        int ty, b0, S;
        uint32_t k = 0, pc0 = 0;
        while (ok && it.next(ty, b0, S)) {
            const int nq = (S + 1) >> 1;
            for (int j = 0; ok && j < S + 3; ++j, ++k) {
                const uint32_t slot = k % NSLOT, ph = (k / NSLOT) & 1;
                ok = tc05::mbar_wait(slot_full + slot, ph, err);
                tc05::tc_fence_after();
                const uint32_t so = (slot * SLOT_BYTES) >> 4;
                const int q_lo = j > 4 ? (j - 3) >> 1 : 0, q_hi = min(nq - 1, j >> 1);
                for (int q = q_lo; ok && q <= q_hi; ++q) {
                    const uint32_t pc = pc0 + (uint32_t)q;
                    if (pc % NISS != par) continue;
                    const int d = j - 2 * q, last = 3 + ((2 * q + 1 < S) ? 1 : 0);
                    if (d > last) continue;
                    const uint32_t acc = pc & 3u;
                    if (d == 0) {
                        ok = tc05::mbar_wait(t_empty + acc, ((pc >> 2) & 1u) ^ 1u, err);
                        tc05::tc_fence_after();
                        TRACE(2, pc);
                        if (!ok) break;
                    }
                    TRACE(3, pc * 8 + d);
                    if (tc05::elect_one()) {
                        const uint32_t dt = tmem_base + acc * 128u;
                        const uint32_t wb = w_lo + (uint32_t)((ctc::c1v4_block(0, d) * 2048) >> 4);
#pragma unroll
                        for (int ky = 0; ky < 7; ++ky) {
                            const uint32_t plo = (ky % 3 == 0 ? p0_lo : p12_lo) + so + (uint32_t)(px_off(ky) >> 4);
                            const uint32_t wlo = wb + (uint32_t)((5 * ky * 2048) >> 4);
                            tc05::mma_bf16_lh(dt, wlo, w_hi, plo, p_hi, idesc, (d == 0 && ky == 0) ? 0u : 1u);
                        }
                        if (d == last) tc05::mma_commit(t_full + acc);
                    }
                    __syncwarp();
                    TRACE(4, pc * 8 + d);
                }
                if (ok && tc05::elect_one()) tc05::mma_commit(slot_empty + slot);   // arrives once this issuer's MMAs on the plane are done
                __syncwarp();
            }
            pc0 += (uint32_t)nq;
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue: group (warp - 4) / 4 drains the pairs with pc % NEG == group
        const uint32_t grp = (uint32_t)(warp - 4) >> 2;
        const int ew = warp & 3;                 // TMEM lane quadrant: ew>>1 = sample of the pair, ew&1 = channel half
        const int jl = lane & 3;                 // Toeplitz column j of this lane = the pooling window it finishes
        const int co = (ew & 1) * 8 + (lane >> 2);
        const float bias_v = bias[co];
        uint8_t* pw = smem + OFF_P + (warp - 4) * PW_BYTES;
        __nv_bfloat16* pw_v = reinterpret_cast<__nv_bfloat16*>(pw);
        uint8_t* pw_a = pw + 56 * 16;
        // Quad exchange (t = column inside the window): conv column 3*jl + t of the 12-column period lives in lane (3 jl + t) % 4
        // as its group register (3 jl + t) / 4. Seen from the sender, the register wanted at step t is (r0 + t) % 3 with
        // r0 = (3 L) / 4 of the asking lane L = 3 j % 4: every lane rotates its three group values by r0 once per period.
        int srcl[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) srcl[t] = (lane & ~3) | ((3 * jl + t) & 3);
        const int r0 = (3 * ((3 * jl) & 3)) >> 2;
        const bool rot1 = r0 == 1, rot2 = r0 == 2;
        const uint32_t rsh = 2u * (uint32_t)r0;
        RunIter it(B, sliding);
        int ty, b0, S;
        uint32_t pc0 = 0;
        bool ok = true;
        while (ok && it.next(ty, b0, S)) {
            const int nq = (S + 1) >> 1;
            for (int q = 0; ok && q < nq; ++q) {
                const uint32_t pc = pc0 + (uint32_t)q;
                if (pc % NEG != grp) continue;
                const uint32_t acc = pc & 3u;
                TRACE(5, pc);
                ok = tc05::mbar_wait(t_full + acc, (pc >> 2) & 1u, err);
                if (!ok) break;
                tc05::tc_fence_after();
                TRACE(6, pc);
                const int s = 2 * q + (ew >> 1);
                const bool valid = s < S;
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 128u;
                float V[2][21];
                uint32_t DP[2][7];               // per 3-group period: the three vertical winners (2 bits each)
                // vertical 3-max of one pooled row: conv rows = columns o+g, o+21+g, o+42+g of `a`; strict compares: the first maximum wins
                auto vertical = [&](const float* a, int o, float* Vr, uint32_t* Dr) {
#pragma unroll
                    for (int p = 0; p < 7; ++p) {
                        uint32_t d3[3];
#pragma unroll
                        for (int e = 0; e < 3; ++e) {
                            const int g = 3 * p + e;
                            const float x0 = a[o + g], x1 = a[o + NG + g], x2 = a[o + 2 * NG + g];
                            const float m1 = fmaxf(x0, x1);
                            Vr[g] = fmaxf(m1, x2);
                            d3[e] = x2 > m1 ? (2u << (2 * e)) : (x1 > x0 ? (1u << (2 * e)) : 0u);
                        }
                        Dr[p] = d3[0] | d3[1] | d3[2];
                    }
                };
                if (valid) {
                    {   // pooled row 1 first (columns 63..125; the x16 loads start at column 48), then row 0, so that at most 80 + 28 values are live
                        float u[80];
#pragma unroll
                        for (int c0 = 0; c0 < 80; c0 += 16) tc05::tmem_ld16(taddr + 48 + c0, u + c0);
                        tc05::tmem_ld_wait();
                        vertical(u, 15, V[1], DP[1]);
                    }
                    float v[64];
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 16) tc05::tmem_ld16(taddr + c0, v + c0);
                    tc05::tmem_ld_wait();
                    tc05::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc05::mbar_arrive(t_empty + acc);   // accumulator drained (this warp's quadrant): released before the arithmetic
                    TRACE(7, pc);
                    vertical(v, 0, V[0], DP[0]);
                } else {
                    tc05::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc05::mbar_arrive(t_empty + acc);
                    continue;                                          // odd run length: the pair's second half is nobody's sample
                }
                const int b = b0 + s;
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {
                    const size_t row = (((size_t)b * 16 + co) * 28 + 2 * ty + pr) * 28;
#pragma unroll
                    for (int p = 0; p < 7; ++p) {
                        const float g0 = V[pr][3 * p], g1 = V[pr][3 * p + 1], g2 = V[pr][3 * p + 2];
                        // rotate by r0: what this lane sends at step t is its group (r0 + t) % 3, value and winner code alike
                        const float s0 = rot1 ? g1 : rot2 ? g2 : g0, s1 = rot1 ? g2 : rot2 ? g0 : g1, s2 = rot1 ? g0 : rot2 ? g1 : g2;
                        const uint32_t dpp = DP[pr][p];
                        const uint32_t dr = (dpp | (dpp << 6)) >> rsh;                    // code of group (r0 + t) % 3 at bits 2t
                        const float rv0 = __shfl_sync(0xffffffffu, s0, srcl[0]), rv1 = __shfl_sync(0xffffffffu, s1, srcl[1]), rv2 = __shfl_sync(0xffffffffu, s2, srcl[2]);
                        const uint32_t rd0 = __shfl_sync(0xffffffffu, dr, srcl[0]) & 3u, rd1 = (__shfl_sync(0xffffffffu, dr, srcl[1]) >> 2) & 3u,
                                       rd2 = (__shfl_sync(0xffffffffu, dr, srcl[2]) >> 4) & 3u;
                        // row-major first maximum of the 3x3 window: larger value, or equal value in an earlier row (columns ascend with t)
                        float best = rv0; uint32_t bd = rd0, bt = 0;
                        if (rv1 > best || (rv1 == best && rd1 < bd)) { best = rv1; bd = rd1; bt = 1; }
                        if (rv2 > best || (rv2 == best && rd2 < bd)) { best = rv2; bd = rd2; bt = 2; }
                        const uint32_t bi = bd * 3u + bt;
                        const int px = 4 * p + jl;
                        const float o = fmaxf(best + bias_v, 0.f);
                        y[row + px] = o;
                        amax[row + px] = (uint8_t)bi;
                        pw_v[(pr * 28 + px) * 8 + (co & 7)] = __float2bfloat16_rn(o);
                        pw_a[(pr * 28 + px) * 8 + (co & 7)] = (uint8_t)bi;
                    }
                }
                __syncwarp();
                // the bf16 copy conv2's shifted-window kernel reads (P8 = [b][c/8][pixel][8], conv_sw.cu) and the routing in the
                // same order (conv1's wgrad builders): the warp's (sample, 8 channels) tile is 56 consecutive pixels of each
                const size_t t8 = ((size_t)b * 2 + (ew & 1)) * 784 + (size_t)(2 * ty) * 28;
                if (ybf) {
                    reinterpret_cast<uint4*>(ybf)[t8 + lane] = reinterpret_cast<const uint4*>(pw)[lane];
                    if (lane < 24) reinterpret_cast<uint4*>(ybf)[t8 + 32 + lane] = reinterpret_cast<const uint4*>(pw)[32 + lane];
                }
                if (amax_p8 && lane < 28) reinterpret_cast<uint4*>(amax_p8 + t8 * 8)[lane] = reinterpret_cast<const uint4*>(pw_a)[lane];
                __syncwarp();                                      // pw is rewritten by the warp's next tile
                TRACE(8, pc);
            }
            pc0 += (uint32_t)nq;
        }
    }
    TRACE(9, 0);
    TRACE_END;
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace c1f4

BC_TRACE_EXPORT(bc_debug_c1f4_trace)       // debug builds only: not part of the ABI

int bc_conv1_fwd4_launch(const bc_ctx* c, void* stream) {
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(c1f4::conv1_fwd4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c1f4::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "conv1 forward (gen 4): smem opt-in %d B failed: %s", c1f4::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    const int ntiles = c->batch * c1f4::TILES_PER_FRAME;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    bc::launch_pdl(c1f4::conv1_fwd4_kernel, dim3(grid), dim3(c1f4::NTHREADS), c1f4::SMEM_BYTES, (cudaStream_t)stream,
        (const __nv_bfloat16*)c->x_tp, c->x_tp_stride_n, c->x_tp_stride_c, (const uint8_t*)c->w_packed + ctc::kPackC1V4, c->params + a.b[0],
        c->act[0], c->amax[0], (__nv_bfloat16*)c->act_bf16[0], (c->conv_mode & 16) ? c->amax0_p8 : nullptr, c->batch, c->err_flag);
    BC_CUDA_LAUNCH_CHECK("conv1_fwd4_kernel");
    return BC_OK;
}
