// C-ABI glue: error state, device check, arena layout, forward/backward drivers, fused Adam.
#include "bc_common.cuh"
#include <math.h>

namespace bc {

static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace bc

namespace {

// ---- K10: fused multi-tensor Adam over the flat arena ---------------------------------------
// state (doubles, like the Python scalars torch.optim.Adam works with):
//   [0] lr [1] beta1 [2] beta2 [3] eps [4] step [5] grad_scale [6] step_size (out) [7] bc2_sqrt (out)
__global__ void adam_tick_kernel(double* st) {
    bc::pdl_wait();
    bc::pdl_trigger();
    // torch/optim/adam.py (single-tensor path): bias_correction1 = 1 - beta1 ** step;
    // step_size = lr / bias_correction1; bias_correction2_sqrt = sqrt(1 - beta2 ** step) -- all in f64.
    const double step = st[4] + 1.0;
    st[4] = step;
    st[6] = st[0] / (1.0 - pow(st[1], step));
    st[7] = sqrt(1.0 - pow(st[2], step));
}

// tick folded into the update: every CTA derives the step's scalars from st[4] + 1 itself (same f64 formulas as
// adam_tick_kernel), the last CTA to finish publishes them and the new step count. st has 9 doubles here: [8] is the
// CTA counter (as u32). Saves one launch per step.
__device__ __forceinline__ void adam_scalars(const double* st, double* s_sc) {
    const double step = st[4] + 1.0;
    s_sc[0] = step;
    s_sc[1] = st[0] / (1.0 - pow(st[1], step));
    s_sc[2] = sqrt(1.0 - pow(st[2], step));
}
__device__ __forceinline__ void adam_publish(double* st, const double* s_sc) {   // called by thread 0 of every CTA after its work
    __threadfence();
    unsigned int* ctr = reinterpret_cast<unsigned int*>(st + 8);
    if (atomicAdd(ctr, 1u) == gridDim.x - 1) {
        st[4] = s_sc[0]; st[6] = s_sc[1]; st[7] = s_sc[2];
        *ctr = 0u;
    }
}

__global__ void __launch_bounds__(256) adam_tick_step_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                             float4* __restrict__ m, float4* __restrict__ v,
                                                             double* __restrict__ st, int64_t n4) {
    __shared__ double s_sc[3];
    bc::pdl_wait();
    bc::pdl_trigger();
    if (threadIdx.x == 0) adam_scalars(st, s_sc);
    __syncthreads();
    const float w1 = (float)(1.0 - st[1]), b2 = (float)st[2], w2 = (float)(1.0 - st[2]), eps = (float)st[3];
    const float gs = (float)st[5], neg_step = (float)(-s_sc[1]), bc2s = (float)s_sc[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = G[k] * gs;
            M[k] = __fmaf_rn(w1, gk - M[k], M[k]);
            V[k] = __fmaf_rn(w2 * gk, gk, V[k] * b2);
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(V[k]), bc2s), eps);
            P[k] = __fmaf_rn(neg_step, __fdiv_rn(M[k], denom), P[k]);
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
    __syncthreads();
    if (threadIdx.x == 0) adam_publish(st, s_sc);
}

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v,
                                                   const double* __restrict__ st, int64_t n4) {
    bc::pdl_wait();
    bc::pdl_trigger();
    // the f64 scalars are rounded to f32 exactly where ATen rounds its Scalar arguments
    const float w1 = (float)(1.0 - st[1]), b2 = (float)st[2], w2 = (float)(1.0 - st[2]), eps = (float)st[3];
    const float gs = (float)st[5], neg_step = (float)(-st[6]), bc2s = (float)st[7];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = G[k] * gs;                          // grad_scale folds the DDP 1/world mean
            M[k] = __fmaf_rn(w1, gk - M[k], M[k]);               // exp_avg.lerp_(grad, 1 - beta1), weight < 0.5 branch
            V[k] = __fmaf_rn(w2 * gk, gk, V[k] * b2);            // exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(V[k]), bc2s), eps);   // (sqrt(v) / bc2_sqrt).add_(eps)
            P[k] = __fmaf_rn(neg_step, __fdiv_rn(M[k], denom), P[k]);                // addcdiv_(m, denom, value=-step_size)
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

// ---- K11 + K10 fused: the data-parallel gradient exchange inside the Adam step, over NVLink peer memory -------------
// Every rank's gradient arena lives in a symmetric allocation that all ranks map (torch symmetric memory, i.e.
// cudaIpc / fabric handles: plumbing). One kernel per step and rank:
//   A  announce "my gradients of epoch e are complete" in every peer's signal pad, wait for every peer's announcement;
//   B  read all `world` arenas straight from peer memory, add them in RANK ORDER (so every rank computes bitwise the same
//      sum and the replicas never drift), scale by grad_scale = 1/world, apply the Adam update to the local replica;
//   C  the last CTA announces "done reading" and waits for the peers' same flag, so that when this kernel has finished
//      on a rank, nobody is still reading that rank's gradients and the next backward may overwrite them.
// 533 KB per arena: the whole exchange is (world-1) x 533 KB of NVLink reads per rank and two flag round trips -- no
// separate all-reduce launch, no host involvement, CUDA-graph capturable. Waits are bounded by wall-clock time.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ bool wait_epoch(const uint32_t* flag, uint32_t epoch, int* err) {
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        if (globaltimer_ns() - t0 > 20000000000ull) {       // 20 s: a peer died or the ranks disagree on the step count
            if (err) atomicExch(err, 2);
            return false;
        }
        __nanosleep(100);
    }
    return true;
}

constexpr int kSigReady = 256, kSigDone = 320;               // u32 word offsets inside a rank's signal pad (64 ranks each)

__global__ void __launch_bounds__(256) adam_exchange_kernel(float4* __restrict__ p, const float4* const* __restrict__ peer_grads,
                                                            uint32_t* const* __restrict__ peer_signals, float4* __restrict__ m,
                                                            float4* __restrict__ v, double* __restrict__ st,
                                                            uint32_t* __restrict__ sync, int64_t n4, int rank, int world, int* err) {
    __shared__ int s_last;
    __shared__ double s_sc[3];
    bc::pdl_wait();
    bc::pdl_trigger();
    if (threadIdx.x == 0) adam_scalars(st, s_sc);          // the tick is folded in (published by the last CTA, phase C)
    const uint32_t epoch = sync[0] + 1;
    uint32_t* my_sig = peer_signals[rank];
    // ---- A
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peer_signals[threadIdx.x] + kSigReady + rank, epoch);
    }
    if (threadIdx.x < world) wait_epoch(my_sig + kSigReady + threadIdx.x, epoch, err);
    __syncthreads();
    // ---- B
    const float w1 = (float)(1.0 - st[1]), b2 = (float)st[2], w2 = (float)(1.0 - st[2]), eps = (float)st[3];
    const float gs = (float)st[5], neg_step = (float)(-s_sc[1]), bc2s = (float)s_sc[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gg = ld_relaxed_sys_f4(peer_grads[0] + i);
        for (int r = 1; r < world; ++r) {
            const float4 o = ld_relaxed_sys_f4(peer_grads[r] + i);
            gg.x = __fadd_rn(gg.x, o.x); gg.y = __fadd_rn(gg.y, o.y); gg.z = __fadd_rn(gg.z, o.z); gg.w = __fadd_rn(gg.w, o.w);
        }
        float4 pp = p[i], mm = m[i], vv = v[i];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = G[k] * gs;
            M[k] = __fmaf_rn(w1, gk - M[k], M[k]);
            V[k] = __fmaf_rn(w2 * gk, gk, V[k] * b2);
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(V[k]), bc2s), eps);
            P[k] = __fmaf_rn(neg_step, __fdiv_rn(M[k], denom), P[k]);
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
    // ---- C
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(sync + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < world) {
            st_release_sys(peer_signals[threadIdx.x] + kSigDone + rank, epoch);
            wait_epoch(my_sig + kSigDone + threadIdx.x, epoch, err);
        }
        __syncthreads();
        if (threadIdx.x == 0) { sync[1] = 0; sync[0] = epoch; st[4] = s_sc[0]; st[6] = s_sc[1]; st[7] = s_sc[2]; }
    }
}

}  // namespace

extern "C" {

const char* bc_last_error_string(void) { return bc::err_buf(); }
int bc_abi_version(void) { return 1; }

int bc_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) return bc::fail(BC_ERR_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
    return BC_OK;
}

int64_t bc_arena_layout(int obs_size, int n_actions, int64_t offsets[14], int64_t sizes[14]) {
    const bc::Arena a = bc::arena_layout(obs_size, n_actions);
    for (int l = 0; l < 7; ++l) {
        offsets[2 * l] = a.w[l]; sizes[2 * l] = a.nw[l];
        offsets[2 * l + 1] = a.b[l]; sizes[2 * l + 1] = a.nb[l];
    }
    return a.total;
}

size_t bc_partials_floats(int obs_size, int n_actions) {
    return (size_t)bc::partials_layout(bc::arena_layout(obs_size, n_actions)).total;
}

int bc_forward(const bc_ctx* c, void* stream) {
    for (int l = 0; l < 4; ++l) {
        int rc = bc_conv_relu_pool_fwd(c, l, stream);
        if (rc) return rc;
    }
    return bc_head(c, 0, stream);
}

int bc_backward(const bc_ctx* c, int with_loss, void* stream) {
    BC_CHECK_ARG(c, "bc_backward: null ctx");
    // head: (CE from labels when with_loss, else the caller's dlogits) + MLP backward -> ghead
    int rc = bc_head(c, with_loss ? 3 : 2, stream);
    if (rc) return rc;
    for (int l = 3; l >= 1; --l) {
        rc = bc_conv_bwd_wgrad(c, l, stream);
        if (!rc) rc = bc_conv_bwd_dgrad(c, l, stream);
        if (rc) return rc;
    }
    if ((rc = bc_conv_bwd_wgrad(c, 0, stream))) return rc;
    return bc_reduce_partials(c, with_loss, stream);
}

int bc_adam_tick(double* state, void* stream) {
    BC_CHECK_ARG(state, "bc_adam_tick: null state");
    bc::launch_pdl(adam_tick_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, state);
    BC_CUDA_LAUNCH_CHECK("adam_tick_kernel");
    return BC_OK;
}

int bc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const double* state,
                 int64_t n, void* stream) {
    BC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state, "bc_adam_step: null pointer");
    BC_CHECK_ARG(n >= 0 && n % 4 == 0, "bc_adam_step: n=%lld must be a multiple of 4 (the arena is padded)", (long long)n);
    BC_CHECK_ARG(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "bc_adam_step: 16 B alignment");
    if (n == 0) return BC_OK;
    const int64_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    const int cap = bc::num_sms() * 8;
    if (blocks > cap) blocks = cap;
    bc::launch_pdl(adam_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (float4*)params, (const float4*)grads, (float4*)exp_avg,
                   (float4*)exp_avg_sq, state, n4);
    BC_CUDA_LAUNCH_CHECK("adam_kernel");
    return BC_OK;
}

int bc_adam_tick_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, double* state9, int64_t n, void* stream) {
    BC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state9, "bc_adam_tick_step: null pointer");
    BC_CHECK_ARG(n >= 0 && n % 4 == 0, "bc_adam_tick_step: n=%lld must be a multiple of 4 (the arena is padded)", (long long)n);
    BC_CHECK_ARG(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "bc_adam_tick_step: 16 B alignment");
    const int64_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    const int cap = bc::num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;                               // n == 0 still ticks
    bc::launch_pdl(adam_tick_step_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (float4*)params, (const float4*)grads,
                   (float4*)exp_avg, (float4*)exp_avg_sq, state9, n4);
    BC_CUDA_LAUNCH_CHECK("adam_tick_step_kernel");
    return BC_OK;
}

int bc_adam_step_exchange(float* params, const void* peer_grads_dev, const void* peer_signals_dev, float* exp_avg, float* exp_avg_sq,
                          double* state, uint32_t* sync_state, int64_t n, int rank, int world, int* err_flag, void* stream) {
    BC_CHECK_ARG(params && peer_grads_dev && peer_signals_dev && exp_avg && exp_avg_sq && state && sync_state, "bc_adam_step_exchange: null pointer");
    BC_CHECK_ARG(n > 0 && n % 4 == 0, "bc_adam_step_exchange: n=%lld must be a positive multiple of 4", (long long)n);
    BC_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bc_adam_step_exchange: rank %d of %d", rank, world);
    const int64_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 64) blocks = 64;                              // every CTA spins on the flags: keep them all resident
    bc::launch_pdl(adam_exchange_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (float4*)params, (const float4* const*)peer_grads_dev,
                   (uint32_t* const*)peer_signals_dev, (float4*)exp_avg, (float4*)exp_avg_sq, state, sync_state, n4, rank, world, err_flag);
    BC_CUDA_LAUNCH_CHECK("adam_exchange_kernel");
    return BC_OK;
}

}  // extern "C"
