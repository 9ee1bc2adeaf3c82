// K13: the tail of the policy forward for closed-loop serving at small batch, ONE launch:
//   conv4x4(32->64)+ReLU+pool2 -> conv3x3(64->128)+ReLU+pool2 -> flatten -> fc 128->64->32->n_actions -> greedy action
// = ConvNet1.cnn_base[6:12] + fc (/root/reference/src/architectures/nets.py:24-33,35-39) and the argmax of
// Imitation.forward's logits the rollout code takes (/root/reference/src/data/stat.py:41). At batch 1..16 the forward is a
// chain of launch latencies (BASELINE configs[4]: 7 dependent launches = 27.5 us at B = 1 with every launch under programmatic
// dependent launch); conv3, conv4, the head and the argmax are 5.9 MFLOP per sample -- this kernel replaces those four
// launches by one.
//
// One thread-block CLUSTER of 8 CTAs per sample, exact f32 FFMA arithmetic on the f32 master weights (both modes):
//   phase A (before griddepcontrol.wait: the parameters are written only by launches that release their dependents after their
//            last write, include/bc_b200.h "Launch ordering contract"): CTA r loads its weight slices into shared memory --
//            conv3 output channels [8r, 8r+8) (16 KB), conv4 output channels [16r, 16r+16) (36 KB), CTA 0 also the head (42 KB);
//   phase B  act2 of the sample (32x12x12 f32, 18 KB) -> shared memory; conv3 for the CTA's 8 channels at the 8x8 conv pixels
//            the floor-mode pool keeps (row / column 8 of the 9x9 output is dropped by MaxPool2d(2), nets.py:26): a thread owns
//            (2 channels, one 2x2 pool window, 8 of the 32 input channels) = 1,024 FMA from a 5x5 input patch in registers;
//            the four input-channel quarters meet in shared memory in fixed order, + bias, ReLU, 2x2 max, and every CTA's
//            128 pooled values are written into ALL eight CTAs' shared memory (distributed shared memory);
//   phase C  conv4 for the CTA's 16 channels: a thread owns (channel, 4 of the 64 input channels) for all 2x2 conv pixels,
//            16-lane xor tree, + bias, ReLU, max -> the 16 features go to CTA 0's shared memory;
//   phase D  CTA 0: the three Linear layers with the summation order of head_kernel (csrc/head.cu), logits, first maximum.
// Two cluster barriers per sample; no global-memory round trip between the layers.
#include "bc_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int CL = 8;                 // CTAs per sample = cluster size (portable maximum)
constexpr int NT = 256;
constexpr int MAXA = BC_MAX_ACTIONS;
constexpr int A3P = 20;               // floats between two channels of the pooled conv3 activation in shared memory (16 used):
                                      // 80 B pitch makes the 16 B patch loads of conv4 conflict-free

// shared-memory plan (floats)
constexpr int O_W3 = 0;                       // [8][32][16]
constexpr int O_W4 = O_W3 + 8 * 32 * 16;      // [16][64][9]
constexpr int O_A2 = O_W4 + 16 * 64 * 9;      // [32][12][12]
constexpr int O_PART = O_A2 + 32 * 144;       // [4 quarters][64 threads][8]
constexpr int O_A3 = O_PART + 4 * 64 * 8;     // [64][A3P]
constexpr int O_A4 = O_A3 + 64 * A3P;         // [128]
constexpr int O_B3 = O_A4 + 128;              // [8]
constexpr int O_B4 = O_B3 + 8;                // [16]
constexpr int O_HW0 = O_B4 + 16;              // head, CTA 0 only: fc.0 [64][128]
constexpr int O_HW2 = O_HW0 + 64 * 128;       // fc.2 [32][64]
constexpr int O_HW4 = O_HW2 + 32 * 64;        // fc.4 [MAXA][32]
constexpr int O_HB0 = O_HW4 + MAXA * 32;
constexpr int O_HB2 = O_HB0 + 64;
constexpr int O_HB4 = O_HB2 + 32;
constexpr int O_H1 = O_HB4 + MAXA;
constexpr int O_H2 = O_H1 + 64;
constexpr int O_Z = O_H2 + 32;
constexpr int SMEM_FLOATS = O_Z + MAXA;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
static_assert(O_W4 % 4 == 0 && O_A2 % 4 == 0 && O_PART % 4 == 0 && O_A3 % 4 == 0 && O_HW0 % 4 == 0 && O_HW2 % 4 == 0 && O_HW4 % 4 == 0,
              "16 B alignment of the vector-accessed regions");

struct TailArgs {
    const float* act2;                                  // (B,32,12,12) f32: pooled output of conv2 (bc_ctx.act[1])
    const float* w3; const float* b3;                   // cnn_base.6: (64,32,4,4), (64)
    const float* w4; const float* b4;                   // cnn_base.9: (128,64,3,3), (128)
    const float* f0; const float* g0; const float* f2; const float* g2; const float* f4; const float* g4;   // fc.{0,2,4} weight, bias
    float* act3; float* act4; float* hid1; float* hid2; // optional outputs (B,64,4,4) (B,128) (B,64) (B,32)
    float* logits; int64_t* actions;                    // (B,NA), (B)
    int B, NA;
};

__global__ void __launch_bounds__(NT, 1) policy_tail_kernel(const TailArgs a) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NA = a.NA;

    // every CTA of the cluster has started once this barrier completes (waited on before the first remote store)
    (void)cluster.barrier_arrive();

    // ---- phase A: weight slices (L2 loads: an L1 line of an earlier launch on this SM could be stale) -----------------
    {
        const float4* src3 = reinterpret_cast<const float4*>(a.w3 + (size_t)r * (8 * 32 * 16));
        float4 t3[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) t3[q] = __ldcg(src3 + tid + NT * q);
        const float4* src4 = reinterpret_cast<const float4*>(a.w4 + (size_t)r * (16 * 64 * 9));
        float4 t4[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) t4[q] = __ldcg(src4 + tid + NT * q);
#pragma unroll
        for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(sm + O_W3)[tid + NT * q] = t3[q];
#pragma unroll
        for (int q = 0; q < 9; ++q) reinterpret_cast<float4*>(sm + O_W4)[tid + NT * q] = t4[q];
        if (tid < 8) sm[O_B3 + tid] = __ldcg(a.b3 + 8 * r + tid);
        if (tid < 16) sm[O_B4 + tid] = __ldcg(a.b4 + 16 * r + tid);
    }
    if (r == 0) {
        const float4* w0v = reinterpret_cast<const float4*>(a.f0);
        float4 t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) t[q] = __ldcg(w0v + tid + NT * q);
        const float4* w2v = reinterpret_cast<const float4*>(a.f2);
        const float4 u0 = __ldcg(w2v + tid), u1 = __ldcg(w2v + tid + NT);
        const bool has4 = tid < NA * 8;
        const float4 u4 = has4 ? __ldcg(reinterpret_cast<const float4*>(a.f4) + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(sm + O_HW0)[tid + NT * q] = t[q];
        reinterpret_cast<float4*>(sm + O_HW2)[tid] = u0;
        reinterpret_cast<float4*>(sm + O_HW2)[tid + NT] = u1;
        if (has4) reinterpret_cast<float4*>(sm + O_HW4)[tid] = u4;
        if (tid < 64) sm[O_HB0 + tid] = __ldcg(a.g0 + tid);
        if (tid < 32) sm[O_HB2 + tid] = __ldcg(a.g2 + tid);
        if (tid < NA) sm[O_HB4 + tid] = __ldcg(a.g4 + tid);
    }
    bc::pdl_wait();                      // act2 is the previous launch's output
    bc::pdl_trigger();

    // ---- phase B: conv3 + ReLU + pool -------------------------------------------------------------------------------------
    {
        const float4* src = reinterpret_cast<const float4*>(a.act2 + (size_t)b * (32 * 144));
        float4 t[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) { const int i = tid + NT * q; t[q] = i < 1152 ? __ldcg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int q = 0; q < 5; ++q) { const int i = tid + NT * q; if (i < 1152) reinterpret_cast<float4*>(sm + O_A2)[i] = t[q]; }
    }
    __syncthreads();
    {
        const int p = tid & 15, py = p >> 2, px = p & 3;     // pooled pixel = 2x2 window of conv pixels (2py.., 2px..)
        const int cp = (tid >> 4) & 3;                       // channel pair: local channels 2cp, 2cp + 1
        const int ciq = tid >> 6;                            // input channels [8 ciq, 8 ciq + 8)
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
            const int ci = ciq * 8 + i;
            const float* ap = sm + O_A2 + ci * 144 + (2 * py) * 12 + 2 * px;
            float pt[5][5];
#pragma unroll
            for (int rr = 0; rr < 5; ++rr) {
                const float2 u = *reinterpret_cast<const float2*>(ap + rr * 12);
                const float2 v = *reinterpret_cast<const float2*>(ap + rr * 12 + 2);
                pt[rr][0] = u.x; pt[rr][1] = u.y; pt[rr][2] = v.x; pt[rr][3] = v.y; pt[rr][4] = ap[rr * 12 + 4];
            }
            const float4* wa = reinterpret_cast<const float4*>(sm + O_W3 + ((2 * cp) * 32 + ci) * 16);
            const float4* wb = reinterpret_cast<const float4*>(sm + O_W3 + ((2 * cp + 1) * 32 + ci) * 16);
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                const float4 w0 = wa[ky], w1 = wb[ky];
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        const float* row = pt[dy + ky];
                        float s0 = acc0[dy * 2 + dx], s1 = acc1[dy * 2 + dx];
                        s0 = fmaf(w0.x, row[dx], s0); s0 = fmaf(w0.y, row[dx + 1], s0); s0 = fmaf(w0.z, row[dx + 2], s0); s0 = fmaf(w0.w, row[dx + 3], s0);
                        s1 = fmaf(w1.x, row[dx], s1); s1 = fmaf(w1.y, row[dx + 1], s1); s1 = fmaf(w1.z, row[dx + 2], s1); s1 = fmaf(w1.w, row[dx + 3], s1);
                        acc0[dy * 2 + dx] = s0; acc1[dy * 2 + dx] = s1;
                    }
                }
            }
        }
        float4* part = reinterpret_cast<float4*>(sm + O_PART + (ciq * 64 + (tid & 63)) * 8);
        part[0] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
        part[1] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
    }
    __syncthreads();
    cluster.barrier_wait();              // all eight CTAs run: their shared memory may be written
    if (tid < 128) {
        const int q = tid & 63, h = tid >> 6;                 // q = (cp, p) of the compute threads, h = channel inside the pair
        const int p = q & 15, cl = 2 * (q >> 4) + h;          // local channel 0..7
        float4 s = *reinterpret_cast<const float4*>(sm + O_PART + q * 8 + h * 4);
#pragma unroll
        for (int k = 1; k < 4; ++k) {                         // input-channel quarters in fixed order
            const float4 t = *reinterpret_cast<const float4*>(sm + O_PART + (k * 64 + q) * 8 + h * 4);
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        const float v = fmaxf(fmaxf(fmaxf(s.x, s.y), fmaxf(s.z, s.w)) + sm[O_B3 + cl], 0.f);   // max_k relu(s_k + bias) = relu(max_k s_k + bias)
        const int cg_ = 8 * r + cl;                           // channel of the 64
#pragma unroll
        for (int d = 0; d < CL; ++d) cluster.map_shared_rank(sm + O_A3, d)[cg_ * A3P + p] = v;
        if (a.act3) a.act3[((size_t)b * 64 + cg_) * 16 + p] = v;
    }
    cluster.sync();                      // the 64 x 4 x 4 activation is complete in every CTA

    // ---- phase C: conv4 + ReLU + pool -------------------------------------------------------------------------------------
    {
        const int s = tid & 15, c = tid >> 4;                 // input channels s, s+16, s+32, s+48 of local output channel c
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ci = s + 16 * i;
            const float4* ip = reinterpret_cast<const float4*>(sm + O_A3 + ci * A3P);
            const float4 r0 = ip[0], r1 = ip[1], r2 = ip[2], r3 = ip[3];
            const float in[4][4] = {{r0.x, r0.y, r0.z, r0.w}, {r1.x, r1.y, r1.z, r1.w}, {r2.x, r2.y, r2.z, r2.w}, {r3.x, r3.y, r3.z, r3.w}};
            const float* wp = sm + O_W4 + (c * 64 + ci) * 9;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float w = wp[ky * 3 + kx];
                    o[0] = fmaf(w, in[ky][kx], o[0]);
                    o[1] = fmaf(w, in[ky][kx + 1], o[1]);
                    o[2] = fmaf(w, in[ky + 1][kx], o[2]);
                    o[3] = fmaf(w, in[ky + 1][kx + 1], o[3]);
                }
            }
        }
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) {               // the 16 input-channel slices of a channel sit in one half-warp
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] += __shfl_xor_sync(0xffffffffu, o[k], off);
        }
        if (s == 0) {
            const float v = fmaxf(fmaxf(fmaxf(o[0], o[1]), fmaxf(o[2], o[3])) + sm[O_B4 + c], 0.f);
            cluster.map_shared_rank(sm + O_A4, 0)[16 * r + c] = v;
            if (a.act4) a.act4[(size_t)b * 128 + 16 * r + c] = v;
        }
    }
    cluster.sync();                      // the 128 features are complete in CTA 0; nobody touches another CTA's memory afterwards
    if (r != 0) return;

    // ---- phase D: fc head + greedy action (summation order of head_kernel) --------------------------------------------------
    const float* s_a = sm + O_A4;
    float* s_h1 = sm + O_H1; float* s_h2 = sm + O_H2; float* s_z = sm + O_Z;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int j = warp * 8 + q;
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) p = fmaf(sm[O_HW0 + j * 128 + lane + 32 * i], s_a[lane + 32 * i], p);
        p = bc::warp_sum(p);
        if (lane == 0) s_h1[j] = fmaxf(p + sm[O_HB0 + j], 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = warp * 4 + q;
        float p = fmaf(sm[O_HW2 + j * 64 + lane], s_h1[lane], sm[O_HW2 + j * 64 + lane + 32] * s_h1[lane + 32]);
        p = bc::warp_sum(p);
        if (lane == 0) s_h2[j] = fmaxf(p + sm[O_HB2 + j], 0.f);
    }
    __syncthreads();
    for (int c = warp; c < NA; c += 8) {
        const float p = bc::warp_sum(sm[O_HW4 + c * 32 + lane] * s_h2[lane]);
        if (lane == 0) s_z[c] = p + sm[O_HB4 + c];
    }
    __syncthreads();
    if (a.hid1 && tid < 64) a.hid1[(size_t)b * 64 + tid] = s_h1[tid];
    if (a.hid2 && tid < 32) a.hid2[(size_t)b * 32 + tid] = s_h2[tid];
    if (tid < NA) a.logits[(size_t)b * NA + tid] = s_z[tid];
    if (warp == 0 && a.actions) {
        const float z = lane < NA ? s_z[lane] : -INFINITY;
        const float m = bc::warp_max(z);
        const unsigned hit = __ballot_sync(0xffffffffu, lane < NA && z == m);   // first maximum, like torch.argmax
        if (lane == 0) a.actions[b] = hit ? (int64_t)(__ffs((int)hit) - 1) : 0;
    }
}

}  // namespace

extern "C" int bc_policy_tail(const bc_ctx* c, int64_t* actions, void* stream) {
    BC_CHECK_ARG(c && c->params && c->act[1] && c->logits, "bc_policy_tail: null buffer (needs params, act[1], logits)");
    BC_CHECK_ARG(c->n_actions >= 1 && c->n_actions <= MAXA, "bc_policy_tail: n_actions %d outside 1..%d", c->n_actions, MAXA);
    BC_CHECK_ARG(c->batch >= 0 && c->batch <= BC_POLICY_TAIL_MAX_BATCH, "bc_policy_tail: batch %d outside 0..%d (one 8-CTA cluster per sample; "
                 "larger batches belong to the tensor-core kernels: bc_forward_act)", c->batch, BC_POLICY_TAIL_MAX_BATCH);
    if (c->batch == 0) return BC_OK;
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(policy_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "policy_tail_kernel: smem opt-in %d B failed: %s", SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    const bc::Arena ar = bc::arena_layout(c->obs_size, c->n_actions);
    TailArgs a{};
    a.act2 = c->act[1];
    a.w3 = c->params + ar.w[2]; a.b3 = c->params + ar.b[2];
    a.w4 = c->params + ar.w[3]; a.b4 = c->params + ar.b[3];
    a.f0 = c->params + ar.w[4]; a.g0 = c->params + ar.b[4];
    a.f2 = c->params + ar.w[5]; a.g2 = c->params + ar.b[5];
    a.f4 = c->params + ar.w[6]; a.g4 = c->params + ar.b[6];
    a.act3 = c->act[2]; a.act4 = c->act[3]; a.hid1 = c->hid1; a.hid2 = c->hid2;
    a.logits = c->logits; a.actions = actions;
    a.B = c->batch; a.NA = c->n_actions;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c->batch * CL); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, policy_tail_kernel, a);
    if (e != cudaSuccess) bc::pending_launch_error() = e;
    BC_CUDA_LAUNCH_CHECK("policy_tail_kernel");
    return BC_OK;
}
