// Self-test of the tcgen05 building blocks (descriptors, TMEM mapping, commit/mbarrier protocol):
// D[M,N] f32 = A[M,K] bf16 (K contiguous) * B[N,K]^T bf16, one CTA per 128-row tile, operands
// staged by the threads into the no-swizzle K-major canonical layout. Exposed through the C ABI so
// the GPU tests can pin the primitives the conv kernels are built from before trusting those.
#include "bc_common.cuh"
#include "tc05.cuh"

namespace {

// canonical no-swizzle K-major tile: 16-byte K-chunk c of row r lives at c*(ROWS*16) + r*16
// => core matrices (8 rows x 16 B) are contiguous 128 B; SBO = 128 B, LBO = ROWS*16 B
template <int NT>
__device__ __forceinline__ void stage_kmajor(uint8_t* smem, const __nv_bfloat16* g, int rows_valid, int ROWS, int K, int ld) {
    const int chunks = K / 8;
    for (int i = threadIdx.x; i < ROWS * chunks; i += NT) {
        const int r = i % ROWS, c = i / ROWS;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < rows_valid) v = *reinterpret_cast<const uint4*>(g + (size_t)r * ld + 8 * c);
        *reinterpret_cast<uint4*>(smem + (size_t)c * ROWS * 16 + r * 16) = v;
    }
}

__global__ void __launch_bounds__(128) tc_gemm_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                               float* __restrict__ D, int M, int N, int K, int* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)128 * K * 2;
    const int m0 = blockIdx.x * 128;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { tc05::mbar_init(&bar, 1); tc05::mbar_fence_init(); }
    if (warp == 0) tc05::tmem_alloc(&tmem_base_s, 256);
    stage_kmajor<128>(sA, A + (size_t)m0 * K, M - m0 < 128 ? M - m0 : 128, 128, K, K);
    stage_kmajor<128>(sB, B, N, N, K, K);
    tc05::fence_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0 && tc05::elect_one()) {
        const uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, 128, (uint32_t)N, 0, 0);
        const uint32_t a0 = tc05::smem_u32(sA), b0 = tc05::smem_u32(sB);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ad = tc05::smem_desc(a0 + ks * 2 * 128 * 16, 128 * 16, 128, tc05::SW_NONE);
            const uint64_t bd = tc05::smem_desc(b0 + ks * 2 * N * 16, N * 16, 128, tc05::SW_NONE);
            tc05::mma_bf16(tmem_base, ad, bd, idesc, ks > 0);
        }
        tc05::mma_commit(&bar);
    }
    __syncwarp();
    const bool ok = tc05::mbar_wait(&bar, 0, err);
    tc05::tc_fence_after();
    if (ok) {
        const int row = m0 + warp * 32 + (threadIdx.x & 31);
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tc05::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
            tc05::tmem_ld_wait();
            if (row < M)
                for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, 256);
}

// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, operands resident
// in shared memory, `reps` back-to-back instructions into one accumulator.
// mode 0: same A/B every time, one commit at the end; mode 1: A cycles over 8 stages;
// mode 2: one tcgen05.commit per MMA (nobody waits on them);
// mode 3+: full producer/consumer ring with NS = mode-2 stages: warp 1+st "refills" stage st (waits the
//          stage's empty barrier, fence.proxy.async, arrives on its full barrier), the MMA thread waits full,
//          issues, commits to empty -- the synchronisation skeleton of the conv kernels without any data movement.
__global__ void __launch_bounds__(512) tc_mma_bench_kernel(int N, int reps, int mode, int M, int mn_major, int a_off16, int a_sbo, long long* cycles, int* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, full[14], empty[14];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NS = mode >= 3 ? mode - 2 : 0;
    for (int i = threadIdx.x; i < (8 * 4096 + 16384 + 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        tc05::mbar_init(&bar, 1);
        for (int i = 0; i < 14; ++i) { tc05::mbar_init(full + i, 1); tc05::mbar_init(empty + i, 1); }
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(&tmem_base_s, 256);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
        // whole warp runs the loop, one elected lane issues (the pattern of the conv kernels)
        // mn_major: both operands MN-major, [core along M/N at 2048 B][16 K rows x 16 B] (the wgrad kernels' layout)
        const uint32_t idesc = tc05::instr_desc(tc05::FMT_BF16, (uint32_t)M, (uint32_t)N, mn_major, mn_major);
        // a_off16 / a_sbo: A operand start shifted by a_off16 x 16 B and an explicit stride between its 8-row core groups -- does a core
        // matrix (128 B) that is not 128 B aligned cost a second shared-memory wavefront?
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem) + 16u * (uint32_t)a_off16, 128, a_sbo ? (uint32_t)a_sbo : (mn_major ? 2048u : 256u), tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + 8 * 4096), 128, mn_major ? 2048 : 256, tc05::SW_NONE);
        const long long t0 = clock64();
        bool ok = true;
        uint32_t st = 0, ph = 0;
        if (mode == 1) {
            // unrolled x8: every MMA of a group has its own uniform operand registers
            for (int i = 0; i < reps; i += 8) {
                if (tc05::elect_one()) {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        tc05::mma_bf16(tmem_base, ad0 + (uint64_t)(mn_major ? u * 16 : u * 256), bd0 + (uint64_t)(u * 16), idesc, (i | u) > 0);
                }
                __syncwarp();
            }
        } else
        for (int i = 0; ok && i < reps; ++i) {
            if (NS) {
                ok = tc05::mbar_wait(full + st, ph, err);
                tc05::tc_fence_after();
            }
            if (tc05::elect_one()) {
                tc05::mma_bf16(tmem_base, ad0 + (mode ? (uint64_t)((i & 7) * 256) : 0), bd0, idesc, i > 0);
                if (NS) tc05::mma_commit(empty + st);
                else if (mode == 2) tc05::mma_commit(full + (i % 14));
            }
            __syncwarp();
            if (NS && ++st == (uint32_t)NS) { st = 0; ph ^= 1; }
        }
        if (tc05::elect_one()) tc05::mma_commit(&bar);
        __syncwarp();
        const long long t1 = clock64();
        tc05::mbar_wait(&bar, 0, err);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
    } else if (NS && warp >= 1 && warp <= NS) {
        const int st = warp - 1;   // warp 0 issues; warps 1..NS are the stand-in producers
        bool ok = true;
        for (int u = 0; ok && st + u * NS < reps; ++u) {
            ok = tc05::mbar_wait(empty + st, (u & 1) ^ 1, err);
            // stand-in for the repack: one 16-byte store per lane into the stage
            reinterpret_cast<uint4*>(smem + (st & 7) * 4096)[lane] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(full + st);
        }
    }
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, 256);
}

}  // namespace

extern "C" int bc_tc_mma_bench(int N, int reps, int mode, int grid, long long* cycles2, int* err_flag, void* stream) {
    const int M = (mode & (1 << 20)) ? 64 : 128;          // bit 20: M=64 instructions
    const int mn_major = (mode >> 21) & 1;                // bit 21: MN-major operands
    const int a_off16 = (mode >> 22) & 7;                 // bits 22-24: A start offset in 16 B units
    const int a_sbo = ((mode >> 25) & 1) ? 2016 : 0;      // bit 25: A core-group stride 2016 B (the un-padded conv1 wgrad slices)
    mode &= ~(0x3f << 20);
    BC_CHECK_ARG(!mn_major || N <= 64, "bc_tc_mma_bench: the MN-major variant has room for N <= 64");
    const int threads = (mode >> 8) ? (mode >> 8) : 512;   // bits 8.. of mode: CTA size override
    mode &= 0xff;
    BC_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && reps > 0 && cycles2 && err_flag, "bc_tc_mma_bench: bad arguments");
    const int smem = 8 * 4096 + 16384 + 1024;
    cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_mma_bench_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(N, reps, mode, M, mn_major, a_off16, a_sbo, cycles2, err_flag);
    BC_CUDA_LAUNCH_CHECK("tc_mma_bench_kernel");
    return BC_OK;
}

extern "C" int bc_tc_gemm_selftest(const void* A, const void* B, float* D, int M, int N, int K, int* err_flag, void* stream) {
    BC_CHECK_ARG(A && B && D && err_flag, "bc_tc_gemm_selftest: null pointer");
    BC_CHECK_ARG(M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K % 16 == 0, "bc_tc_gemm_selftest: need N%%16==0 in [16,256], K%%16==0");
    const size_t smem = (size_t)(128 + N) * K * 2;
    BC_CHECK_ARG(smem <= 200 * 1024, "bc_tc_gemm_selftest: K too large for the single-stage test (%zu B smem)", smem);
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "tc_gemm_selftest: smem opt-in failed: %s", cudaGetErrorString(e));
    tc_gemm_selftest_kernel<<<(M + 127) / 128, 128, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, M, N, K, err_flag);
    BC_CUDA_LAUNCH_CHECK("tc_gemm_selftest_kernel");
    return BC_OK;
}
