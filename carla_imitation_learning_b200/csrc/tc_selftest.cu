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

}  // namespace

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
