// Weight operand images of the tcgen05 conv kernels and the dispatch of conv2-4 (bf16 mode).
// The kernels themselves: conv_sw.cu (conv2, conv3: shifted-window forward / dgrad / wgrad) and conv4_sw.cu.
#include "bc_common.cuh"
#include "tc05.cuh"
#include "pack.cuh"

using ctc::kPackOff1; using ctc::kPackOff2; using ctc::kPackOff3; using ctc::kPackOff4;
using ctc::kPackD2; using ctc::kPackD3; using ctc::kPackD4; using ctc::kPackTotal;

size_t bc_conv_tc_pack_total() { return kPackTotal; }

namespace ctc {
struct PackArgs { const float* w1; const float* w2; const float* w3; const float* w4; uint8_t* base; int obs; };

constexpr int kNW2 = L2::COUT * L2::CIN * L2::KS * L2::KS, kNW3 = L3::COUT * L3::CIN * L3::KS * L3::KS, kNW4 = L4::COUT * L4::CIN * L4::KS * L4::KS;
constexpr int kNC1 = 28 * 64 * 16;                         // conv1's Toeplitz image, element-ordered (it has structural zeros)
constexpr int kNC1V4 = kC1V4Blocks * 1024;                 // the swapped-role image of conv1 (conv1_fwd4.cu), element-ordered as well
constexpr int kPackElems = kNC1 + kNC1V4 + kNW2 + kNW3 + kNW4;
// all seven operand images (conv1 Toeplitz, conv2-4 forward + dgrad) in one launch
// (the dependents are released only after the last write: the conv kernels fetch these images before their own wait)
__device__ __forceinline__ void pack_all_body(const PackArgs& a, size_t o2, size_t o3, size_t o4, size_t d2, size_t d3, size_t d4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int ncam = a.obs == 12 ? 3 : 1;
    if (i < ncam * kNC1) {
        // step s = (ky,ci): 64 rows n = (j*16+co) x 16 k; Wt[(j,co)][(ci,ky,p)] = W[co][ci][ky][p - 3j] or 0 (conv1_tc.cu);
        // obs_size 12: one such image per camera, its channel ci = frame is the network's channel 3*frame + cam
        const int cam = i / kNC1;
        i -= cam * kNC1;
        const int k = i & 15, n = (i >> 4) & 63, st = i >> 10;
        const int ky = st >> 2, ci = st & 3, j = n >> 4, co = n & 15;     // steps ordered [ky][ci]: consecutive channels = consecutive B rows
        const int kx = k - 3 * j;
        const int cnet = a.obs == 12 ? 3 * ci + cam : ci;
        const float v = (kx >= 0 && kx < 7) ? a.w1[((co * a.obs + cnet) * 7 + ky) * 7 + kx] : 0.f;
        reinterpret_cast<__nv_bfloat16*>(a.base + c1_cam_off(cam))[(size_t)st * 1024 + op_off(n, k >> 3) / 2 + (k & 7)] = __float2bfloat16_rn(v);
        return;
    }
    i -= ncam * kNC1;
    if (a.obs == 12) i += kNC1V4;          // the swapped-role image exists for obs_size 4 only: skip its index range
    if (i < kNC1V4) {
        // block (ky, ci) = 1 + 5 ky + (3 - ci), zero blocks at 5 ky; row = co*4 + j; Wt[(co,j)][p] = W[co][ci][ky][p - 3j] or 0
        const int k = i & 15, r = (i >> 4) & 63, blk = i >> 10;
        float v = 0.f;
        if (blk % 5 != 0) {
            const int ky = (blk - 1) / 5, ci = 3 - (blk - 1) % 5, co = r >> 2, j = r & 3, kx = k - 3 * j;
            if (kx >= 0 && kx < 7) v = a.w1[((co * 4 + ci) * 7 + ky) * 7 + kx];
        }
        reinterpret_cast<__nv_bfloat16*>(a.base + kPackC1V4)[(size_t)blk * 1024 + op_off(r, k >> 3) / 2 + (k & 7)] = __float2bfloat16_rn(v);
        return;
    }
    i -= kNC1V4;
    if (i < kNW2) { pack_src_elem<L2>(a.w2[i], (__nv_bfloat16*)(a.base + o2), (__nv_bfloat16*)(a.base + d2), i); return; } i -= kNW2;
    if (i < kNW3) { pack_src_elem<L3>(a.w3[i], (__nv_bfloat16*)(a.base + o3), (__nv_bfloat16*)(a.base + d3), i); return; } i -= kNW3;
    if (i < kNW4) pack_src_elem<L4>(a.w4[i], (__nv_bfloat16*)(a.base + o4), (__nv_bfloat16*)(a.base + d4), i);
}
__global__ void pack_all_kernel(const PackArgs a, size_t o2, size_t o3, size_t o4, size_t d2, size_t d3, size_t d4) {
    bc::pdl_wait();
    pack_all_body(a, o2, o3, o4, d2, d3, d4);
    __threadfence();
    bc::pdl_trigger();
}
}  // namespace ctc

int bc_conv_tc_pack(const bc_ctx* c, void* stream) {
    const bc::Arena a = bc::arena_layout(c->obs_size, c->n_actions);
    ctc::PackArgs pa{c->params + a.w[0], c->params + a.w[1], c->params + a.w[2], c->params + a.w[3], (uint8_t*)c->w_packed, c->obs_size};
    bc::launch_pdl(ctc::pack_all_kernel, dim3((ctc::kPackElems + 2 * ctc::kNC1 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, pa, kPackOff2, kPackOff3, kPackOff4, kPackD2, kPackD3, kPackD4);
    BC_CUDA_LAUNCH_CHECK("pack_all_kernel");
    return BC_OK;
}

int bc_wgrad_tc_launch(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(layer >= 1 && layer <= 3, "wgrad (tcgen05): layer %d", layer);
    BC_CHECK_ARG(c->err_flag && c->act_bf16[layer - 1] && c->partials && c->act[layer] && c->amax[layer],
                 "conv%d wgrad (tcgen05): null buffer", layer + 1);
    return layer == 3 ? bc_conv4_sw_wgrad_launch(c, stream) : bc_conv_sw_wgrad_launch(c, layer, stream);
}

int bc_dgrad_tc_launch(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(layer >= 1 && layer <= 3, "dgrad (tcgen05): layer %d", layer);
    BC_CHECK_ARG(c->w_packed && c->err_flag && c->gact[layer - 1] && c->act[layer] && c->amax[layer],
                 "conv%d dgrad (tcgen05): null buffer", layer + 1);
    const uint8_t* base = (const uint8_t*)c->w_packed;
    switch (layer) {
    case 1: return bc_conv_sw_dgrad_launch(c, 1, base + kPackD2, stream);
    case 2: return bc_conv_sw_dgrad_launch(c, 2, base + kPackD3, stream);
    default: return bc_conv4_sw_dgrad_launch(c, base + kPackD4, stream);
    }
}

int bc_conv_tc_launch(const bc_ctx* c, int layer, void* stream) {
    BC_CHECK_ARG(layer >= 1 && layer <= 3, "conv (tcgen05): layer %d", layer);
    BC_CHECK_ARG(c->w_packed && c->err_flag && c->act_bf16[layer - 1] && c->act[layer] && c->amax[layer],
                 "conv%d (tcgen05): null buffer (w_packed, err_flag, act_bf16 input, act, amax)", layer + 1);
    const uint8_t* base = (const uint8_t*)c->w_packed;
    switch (layer) {
    case 1: return bc_conv_sw_fwd_launch(c, 1, base + kPackOff2, stream);
    case 2: return bc_conv_sw_fwd_launch(c, 2, base + kPackOff3, stream);
    default: return bc_conv4_sw_fwd_launch(c, base + kPackOff4, stream);
    }
}
