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
// Per CTA (persistent, one per SM), tile = 6 conv rows x 84 columns of one frame (= 2 pooled rows):
//   warp 0      loader   : cp.async.bulk (TMA engine), one 11 KB copy per input plane (22 contiguous rows),
//                          pipelined PER PLANE: plane ci of tile t+1 streams in while planes ci+1.. of tile t repack
//   warp 1      MMA      : one thread issues 28 tcgen05.mma per tile (fully unrolled, static descriptors)
//                          into one of two TMEM accumulators
//   warp 2      TMEM allocator
//   warps 4-7   epilogue : tcgen05.ld -> smem -> 3x3 max / first-max argmax / +bias / ReLU -> global
//   warps 8-14  repack   : warp w = kernel row ky: raw rows -> 128x16 bf16 A chunk (UMMA K-major canonical layout)
// mbarrier rings: plane full/empty (4), A-stage full/empty (14 stages), TMEM full/empty (2). All waits bounded.
#include "bc_common.cuh"
#include "tc05.cuh"

namespace c1tc {

constexpr int NREPACK = 7;               // repack warps: warp w owns kernel row ky = w
constexpr int NTHREADS = (8 + NREPACK) * 32;   // 480
constexpr int ROWS_IN = 22;              // input rows per tile: 3*(6-1)+7
constexpr int NG = 21;                   // groups of 4 output columns per conv row
constexpr int MROWS = 126;               // 6 conv rows x 21 groups (of the MMA's 128)
constexpr int NSTEP = 28;                // K steps (ci, ky) of 16 pixels
constexpr int NST = 7;                   // A stages: stage ky holds the TWO chunks (2cp, ky), (2cp+1, ky) of one fill
constexpr int ROW_BYTES = 512;           // 256 px bf16
constexpr int PLANE_BYTES = ROWS_IN * ROW_BYTES;         // 11264: the tile's rows of one input plane, contiguous in HBM
constexpr int A_CHUNK = 128 * 32;                        // 4096: one K=16 slice of the 128-row A tile
constexpr int A_STAGE = 2 * A_CHUNK;                     // 8192
constexpr int B_STEP = 64 * 32;                          // 2048
constexpr int B_BYTES = NSTEP * B_STEP;                  // 57344
constexpr int S_PITCH = 68;                              // floats; 4-bank skew per row => conflict-free STS.128
constexpr int S_BYTES = MROWS * S_PITCH * 4;             // 34272
constexpr int OFF_B = 0;
constexpr int OFF_RAW = OFF_B + B_BYTES;                 // 4 plane buffers (pipelined per plane, not per tile)
constexpr int OFF_A = OFF_RAW + 4 * PLANE_BYTES;
constexpr int OFF_S = OFF_A + NST * A_STAGE;
constexpr int OFF_BAR = (OFF_S + S_BYTES + 127) / 128 * 128;
constexpr int NBAR = 1 + 4 + 4 + NST + NST + 2 + 2;
constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
constexpr int TILES_PER_FRAME = 14;
constexpr int TMEM_COLS = 128;

// UMMA K-major no-swizzle canonical layout used for both operands: a K=16 slice of `rows` rows is
// stored as 8-row groups of 256 B; inside a group the two 16-byte K-chunks are 128 B apart:
//   byte(r, k) = (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2          => LBO = 128 B, SBO = 256 B
__host__ __device__ constexpr int op_off(int r, int chunk) { return (r >> 3) * 256 + chunk * 128 + (r & 7) * 16; }

// fp32 OIHW conv1 weights -> Toeplitz bf16 operand, already in the smem image the MMA reads:
// step s=(ci,ky): 64 rows n=(j*16+co) x 16 k
__global__ void pack_conv1_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NSTEP * 64 * 16) return;
    const int k = i & 15, n = (i >> 4) & 63, s = i >> 10;
    const int ci = s / 7, ky = s % 7, j = n >> 4, co = n & 15;
    const int kx = k - 3 * j;
    const float v = (kx >= 0 && kx < 7) ? w[((co * 4 + ci) * 7 + ky) * 7 + kx] : 0.f;
    out[(size_t)s * (B_STEP / 2) + op_off(n, k >> 3) / 2 + (k & 7)] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv1_tc_kernel(const __nv_bfloat16* __restrict__ x, int64_t sn, int64_t sc, const __nv_bfloat16* __restrict__ wpk,
                const float* __restrict__ bias, float* __restrict__ y, uint8_t* __restrict__ amax,
                __nv_bfloat16* __restrict__ ybf, int B, int* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* b_full = bars;
    uint64_t* raw_full = bars + 1;                 // [4] one per input plane
    uint64_t* raw_empty = bars + 5;                // [4]
    uint64_t* a_full = bars + 9;                   // [NST]
    uint64_t* a_empty = bars + 9 + NST;            // [NST]
    uint64_t* t_full = bars + 9 + 2 * NST;         // [2]
    uint64_t* t_empty = bars + 11 + 2 * NST;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = B * TILES_PER_FRAME;

    if (threadIdx.x == 0) {
        tc05::mbar_init(b_full, 1);
        for (int i = 0; i < 4; ++i) { tc05::mbar_init(raw_full + i, 1); tc05::mbar_init(raw_empty + i, NREPACK); }
        for (int i = 0; i < 2; ++i) { tc05::mbar_init(t_full + i, 1); tc05::mbar_init(t_empty + i, 4); }
        for (int i = 0; i < NST; ++i) { tc05::mbar_init(a_full + i, 1); tc05::mbar_init(a_empty + i, 1); }
        tc05::mbar_fence_init();
    }
    if (warp == 2) tc05::tmem_alloc(tmem_slot, TMEM_COLS);
    // rows 126,127 of every A stage are never produced: keep them zero
    for (int i = threadIdx.x; i < NST * 2 * 2 * 2 * 4; i += NTHREADS) {
        const int ch = i / 16, rem = i % 16, c = rem / 8, r = 126 + (rem % 8) / 4, q = rem % 4;   // ch = stage*2 + chunk
        reinterpret_cast<uint32_t*>(smem + OFF_A + ch * A_CHUNK + op_off(r, c))[q] = 0u;
    }
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ loader (one lane)
        if (lane == 0) {
            tc05::mbar_expect_tx(b_full, B_BYTES);
            tc05::bulk_g2s(smem + OFF_B, wpk, B_BYTES, b_full);
            int it = 0;
            bool ok = true;
            for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
                const int b = t / TILES_PER_FRAME, ty = t % TILES_PER_FRAME;
                const __nv_bfloat16* src = x + (size_t)b * sn + (size_t)(ty * 18) * 256;
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    // plane buffer ci is free once all repack warps are past plane ci of the previous tile
                    ok = ok && tc05::mbar_wait(raw_empty + ci, (it & 1) ^ 1, err);
                    if (!ok) break;
                    tc05::mbar_expect_tx(raw_full + ci, PLANE_BYTES);
                    tc05::bulk_g2s(smem + OFF_RAW + ci * PLANE_BYTES, src + (size_t)ci * sc, PLANE_BYTES, raw_full + ci);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop (so every operand stays in uniform registers); one elected lane
        // issues. Per tile: 14 stage visits (cp-major, ky-minor), 2 MMAs + 1 commit each.
        constexpr uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, 64, 0, 0);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_A), 128, 256, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_B), 128, 256, tc05::SW_NONE);
        bool ok = tc05::mbar_wait(b_full, 0, err);
        int it = 0;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            ok = tc05::mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1, err);
            tc05::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
            for (int v = 0; v < 14; ++v) {
                const int cp = v / 7, ky = v % 7;            // static after unrolling
                if (ok) ok = tc05::mbar_wait(a_full + ky, (uint32_t)(it * 2 + cp) & 1, err);
                tc05::tc_fence_after();
                if (ok && tc05::elect_one()) {
                    // descriptors differ from the base only in the 14-bit start-address field (>>4)
                    const uint64_t a0 = ad0 + (uint64_t)(ky * (A_STAGE >> 4));
                    tc05::mma_bf16(d_tmem, a0, bd0 + (uint64_t)(((2 * cp) * 7 + ky) * (B_STEP >> 4)), idesc, v > 0);
                    tc05::mma_bf16(d_tmem, a0 + (A_CHUNK >> 4), bd0 + (uint64_t)(((2 * cp + 1) * 7 + ky) * (B_STEP >> 4)), idesc, 1);
                    tc05::mma_commit(a_empty + ky);          // stage reusable once both MMAs have read it
                    if (v == 13) tc05::mma_commit(t_full + acc);   // accumulator complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ epilogue
        const int ew = warp - 4;                 // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
        const int te = threadIdx.x - 128;        // 0..127
        const int r = ew * 32 + lane;            // accumulator row
        float* S = reinterpret_cast<float*>(smem + OFF_S);
        float breg[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) breg[i] = bias[(te + 128 * i) & 15];
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            if (!tc05::mbar_wait(t_full + acc, (it >> 1) & 1, err)) break;
            tc05::tc_fence_after();
            float v[64];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 64 + c0, v + c0);
            tc05::tmem_ld_wait();
            tc05::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(t_empty + acc);   // accumulator drained: the MMA may overwrite it
            if (r < MROWS) {
                float4* dst = reinterpret_cast<float4*>(S + r * S_PITCH);
#pragma unroll
                for (int q = 0; q < 16; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");      // S complete (epilogue warps only)
            const int b = t / TILES_PER_FRAME, ty = t % TILES_PER_FRAME;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int o = te + 128 * i;                     // 2 pooled rows x 28 x 16 = 896 outputs
                const int co = o & 15, px = (o >> 4) % 28, pyl = o / 448;
                // S[(oy_l*21 + ox/4) * S_PITCH + (ox%4)*16 + co], ox = 3*px + dx
                int coff[3];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int ox = 3 * px + dx;
                    coff[dx] = (ox >> 2) * S_PITCH + (ox & 3) * 16 + co;
                }
                const float* s0 = S + (3 * pyl) * NG * S_PITCH;
                float best = s0[coff[0]];
                int idx = 0;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        if (dy == 0 && dx == 0) continue;
                        const float vv = s0[dy * NG * S_PITCH + coff[dx]];
                        if (vv > best) { best = vv; idx = dy * 3 + dx; }     // strict: first maximum wins
                    }
                const size_t g = (((size_t)b * 16 + co) * 28 + 2 * ty + pyl) * 28 + px;
                const float out = fmaxf(best + breg[i], 0.f);
                y[g] = out;
                amax[g] = (uint8_t)idx;
                // NHWC bf16 copy for the tensor-core conv2 (16 lanes = 16 channels = 32 contiguous bytes)
                if (ybf) ybf[(((size_t)b * 28 + 2 * ty + pyl) * 28 + px) * 16 + co] = __float2bfloat16_rn(out);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");      // everyone done reading S
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ repack (A producer); warp <-> ky
        const int ky = warp - 8;
        int src_off[4], dst_off[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = q * 32 + lane;
            src_off[q] = r < MROWS ? (3 * (r / NG) + ky) * ROW_BYTES + 24 * (r % NG) : -1;
            dst_off[q] = op_off(r, 0);
        }
        uint32_t fills = 0;                       // fills of stage ky by this warp: 2 per tile (channel pairs)
        bool ok = true;
        int it = 0;
        for (int t = blockIdx.x; ok && t < ntiles; t += gridDim.x, ++it) {
#pragma unroll
            for (int cp = 0; cp < 2; ++cp, ++fills) {
                ok = ok && tc05::mbar_wait(raw_full + 2 * cp, it & 1, err) && tc05::mbar_wait(raw_full + 2 * cp + 1, it & 1, err);
                ok = ok && tc05::mbar_wait(a_empty + ky, (fills & 1) ^ 1, err);
                if (!ok) break;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint8_t* base = smem + OFF_RAW + (2 * cp + h) * PLANE_BYTES;
                    uint8_t* dst = smem + OFF_A + ky * A_STAGE + h * A_CHUNK;
                    uint2 v[4][4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (src_off[q] >= 0) {
                            const uint2* p = reinterpret_cast<const uint2*>(base + src_off[q]);
                            v[q][0] = p[0]; v[q][1] = p[1]; v[q][2] = p[2]; v[q][3] = p[3];
                        }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (src_off[q] >= 0) {
                            uint8_t* d = dst + dst_off[q];
                            *reinterpret_cast<uint4*>(d) = make_uint4(v[q][0].x, v[q][0].y, v[q][1].x, v[q][1].y);
                            *reinterpret_cast<uint4*>(d + 128) = make_uint4(v[q][2].x, v[q][2].y, v[q][3].x, v[q][3].y);
                        }
                }
                tc05::fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tc05::mbar_arrive(a_full + ky);
                    tc05::mbar_arrive(raw_empty + 2 * cp);       // done with both planes of this pair for this tile
                    tc05::mbar_arrive(raw_empty + 2 * cp + 1);
                }
            }
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc05::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace c1tc

extern "C" int bc_pack_weights(const bc_ctx* c, void* stream) {
    BC_CHECK_ARG(c && c->params && c->w_packed, "bc_pack_weights: null buffer");
    BC_CHECK_ARG(c->obs_size == 4, "bc_pack_weights: the tcgen05 conv1 operand exists for obs_size 4 only");
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    c1tc::pack_conv1_weights_kernel<<<(c1tc::NSTEP * 1024 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        c->params + a.w[0], (__nv_bfloat16*)c->w_packed);
    BC_CUDA_LAUNCH_CHECK("pack_conv1_weights_kernel");
    return bc_conv_tc_pack(c, stream);     // conv2-4 operand images follow conv1's inside w_packed
}

extern "C" size_t bc_packed_weight_bytes(void) { return bc_conv_tc_pack_total(); }

int bc_conv1_tc_launch(const bc_ctx* c, void* stream) {
    BC_CHECK_ARG(c->x && c->w_packed && c->err_flag && c->act[0] && c->amax[0], "conv1 (tcgen05): null buffer (x, w_packed, err_flag, act, amax)");
    BC_CHECK_ARG(c->x_dtype == BC_BF16 && c->obs_size == 4, "conv1 (tcgen05): needs bf16 gray planes and obs_size 4");
    BC_CHECK_ARG(((uintptr_t)c->x % 16 == 0) && (c->x_stride_n * 2) % 16 == 0 && (c->x_stride_c * 2) % 16 == 0 && ((uintptr_t)c->w_packed % 16 == 0),
                 "conv1 (tcgen05): x, its strides and w_packed must be 16 B aligned");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(c1tc::conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c1tc::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "conv1 (tcgen05): smem opt-in %d B failed: %s", c1tc::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    const int ntiles = c->batch * c1tc::TILES_PER_FRAME;
    int grid = bc::num_sms();
    if (grid > ntiles) grid = ntiles;
    c1tc::conv1_tc_kernel<<<grid, c1tc::NTHREADS, c1tc::SMEM_BYTES, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)c->x, c->x_stride_n, c->x_stride_c, (const __nv_bfloat16*)c->w_packed, c->params + a.b[0],
        c->act[0], c->amax[0], (__nv_bfloat16*)c->act_bf16[0], c->batch, c->err_flag);
    BC_CUDA_LAUNCH_CHECK("conv1_tc_kernel");
    return BC_OK;
}
