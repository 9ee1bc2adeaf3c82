// Shared device/host helpers for the BC hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/bc_b200.h"

int bc_conv1_tc_launch(const bc_ctx* c, void* stream);            // conv1_tc.cu
int bc_conv1_wgrad_tc_launch(const bc_ctx* c, void* stream);
int bc_conv1_fwd4_launch(const bc_ctx* c, void* stream);           // conv1_fwd4.cu (swapped-role forward; args checked by bc_conv1_tc_launch)
int bc_conv1_wgrad_tp_grid(const bc_ctx* c);   // CTAs = partial-sum slots the tcgen05 conv1 wgrad writes
int bc_conv_tc_launch(const bc_ctx* c, int layer, void* stream);  // conv_tc.cu (layers 1..3)
int bc_conv_tc_pack(const bc_ctx* c, void* stream);
int bc_dgrad_tc_launch(const bc_ctx* c, int layer, void* stream);
int bc_wgrad_tc_launch(const bc_ctx* c, int layer, void* stream);
int bc_conv_sw_fwd_launch(const bc_ctx* c, int layer, const uint8_t* wpk, void* stream);     // conv_sw.cu (layers 1, 2)
int bc_conv_sw_dgrad_launch(const bc_ctx* c, int layer, const uint8_t* wpk, void* stream);
int bc_conv_sw_wgrad_launch(const bc_ctx* c, int layer, void* stream);
int bc_conv4_sw_fwd_launch(const bc_ctx* c, const uint8_t* wpk, void* stream);               // conv4_sw.cu (layer 3)
int bc_conv4_sw_dgrad_launch(const bc_ctx* c, const uint8_t* wpk, void* stream);
int bc_conv4_sw_wgrad_launch(const bc_ctx* c, void* stream);
size_t bc_conv_tc_pack_total();
int bc_head_launch(const bc_ctx* c, int head_mode, int64_t* actions, void* stream);           // head.cu (bc_head + the optional greedy action)

namespace bc {

// ---- error reporting (thread-local message, C-ABI returns the code) -------------------
char* err_buf();
int fail(int code, const char* fmt, ...);
#define BC_CHECK_ARG(cond, ...) do { if (!(cond)) return bc::fail(BC_ERR_ARG, __VA_ARGS__); } while (0)
// the status of the launch just made: launch_pdl's own return value first, then the runtime's sticky-free last error
// (cudaGetLastError CLEARS it, so one failed launch cannot make every later ABI call report failure)
cudaError_t& pending_launch_error();
#define BC_CUDA_LAUNCH_CHECK(name) do { cudaError_t e_ = bc::pending_launch_error(); bc::pending_launch_error() = cudaSuccess; \
    const cudaError_t g_ = cudaGetLastError(); if (e_ == cudaSuccess) e_ = g_; \
    if (e_ != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "%s: %s", name, cudaGetErrorString(e_)); } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per function AND per device: one flag per device ordinal
struct PerDeviceOnce {
    bool done[64] = {};
    bool& operator()() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0; return done[d]; }
};

int num_sms();

// ---- network geometry (nets.py:17-33) at 256x256 ----------------------------------------
struct LayerGeom { int cin, cout, k, s, p, hin, hc, hp; };
__host__ __device__ constexpr LayerGeom geom(int layer, int obs) {
    return layer == 0 ? LayerGeom{obs, 16, 7, 3, 3, 256, 84, 28}
         : layer == 1 ? LayerGeom{16, 32, 5, 1, 2, 28, 24, 12}
         : layer == 2 ? LayerGeom{32, 64, 4, 1, 2, 12, 9, 4}
         :              LayerGeom{64, 128, 3, 1, 2, 4, 2, 1};
}

// ---- arena layout -------------------------------------------------------------------------
// order in memory: fc.4 fc.2 fc.0 conv4 conv3 conv2 conv1 (weight, bias), each padded to 32 floats
struct Arena {
    int64_t w[7], b[7];     // index 0..3 = conv1..4, 4..6 = fc.0, fc.2, fc.4
    int64_t nw[7], nb[7];
    int64_t seg_off[5], seg_len[5];  // gradient segments: head(fc.4..fc.0), conv4, conv3, conv2, conv1
    int64_t total;
};
__host__ inline int64_t pad32(int64_t n) { return (n + 31) / 32 * 32; }
__host__ inline Arena arena_layout(int obs, int na) {
    Arena a{};
    int64_t off = 0;
    const int fin[3] = {128, 64, 32};
    const int fout[3] = {64, 32, na};
    a.seg_off[0] = 0;
    for (int f = 2; f >= 0; --f) {
        a.w[4 + f] = off; a.nw[4 + f] = (int64_t)fin[f] * fout[f]; off += pad32(a.nw[4 + f]);
        a.b[4 + f] = off; a.nb[4 + f] = fout[f];                    off += pad32(a.nb[4 + f]);
    }
    a.seg_len[0] = off;
    for (int l = 3; l >= 0; --l) {
        LayerGeom g = geom(l, obs);
        a.seg_off[4 - l] = off;
        a.w[l] = off; a.nw[l] = (int64_t)g.cout * g.cin * g.k * g.k; off += pad32(a.nw[l]);
        a.b[l] = off; a.nb[l] = g.cout;                               off += pad32(a.nb[l]);
        a.seg_len[4 - l] = off - a.seg_off[4 - l];
    }
    a.total = off;
    return a;
}

// partial-sum workspace: per segment, NPART[seg] copies of seg_len floats, then loss partials
// 128 = 2 samples per head CTA at B = 256. Measured against 256 (one sample per CTA, two CTAs per SM; same box, three alternating runs each,
// profiles/r2M_ab_head_blocks.log): the head kernel alone drops from 11.0 to 7.4 us, but the STEP rises from 0.1911 to 0.1931 ms -- twice the partial
// copies for the reduction to read cold and 256 x 42 KB of weight prologue under conv4's forward cost more than the head saves.
#ifndef BC_HEAD_BLOCKS
#define BC_HEAD_BLOCKS 128
#endif
constexpr int kHeadBlocks = BC_HEAD_BLOCKS;    // partial copies of the head segment (= head CTAs)
constexpr int kWgradParts[4] = {296, 29, 37, 16};  // conv1..conv4: partial-sum slots = grid.x of the wgrad kernels (conv2: x 5 kernel rows = 145 CTAs, conv3: x 4 = 148, conv4: x 6 classes = 96)
struct Partials { int64_t off[5]; int nparts[5]; int64_t loss_off; int64_t total; };
__host__ inline Partials partials_layout(const Arena& a) {
    Partials p{};
    int64_t off = 0;
    p.off[0] = 0; p.nparts[0] = kHeadBlocks; off += a.seg_len[0] * kHeadBlocks;
    for (int s = 1; s < 5; ++s) {  // s=1 -> conv4 ... s=4 -> conv1
        int layer = 4 - s;
        p.off[s] = off; p.nparts[s] = kWgradParts[layer]; off += a.seg_len[s] * kWgradParts[layer];
    }
    p.loss_off = off; off += 32 * ((kHeadBlocks + 31) / 32);
    p.total = off;
    return p;
}

// ---- launch with programmatic stream serialization (see tc05::pdl_wait); the kernel MUST call pdl_wait() before it
// reads or writes anything another kernel of the stream produces or consumes
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
    if (e != cudaSuccess) pending_launch_error() = e;
    return e;
}

// device side of launch_pdl (same instructions as tc05::pdl_wait / pdl_trigger, for kernels that do not use tc05.cuh)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- small device helpers -----------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace bc

int bc_conv1_wgrad3_launch(const bc_ctx* c, const bc::Arena& ar, const bc::Partials& pl, int grid, void* stream);   // conv1_wgrad3.cu
