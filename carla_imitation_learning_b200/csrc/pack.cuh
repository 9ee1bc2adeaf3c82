// bf16 MMA operand images of the tcgen05 conv kernels (bf16 mode): layout constants and the per-weight scatter.
// Shared by conv_tc.cu (the full pack launch: structural zeros of conv1's Toeplitz image + every weight) and
// abi.cu (the Adam kernels refresh the images of the weights they have just updated, so that the step needs no
// separate pack launch).
#pragma once
#include "bc_common.cuh"

namespace ctc {

// UMMA K-major no-swizzle canonical layout of a K=16 slice: byte(r, k) = (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2
__host__ __device__ constexpr int op_off(int r, int chunk) { return (r >> 3) * 256 + chunk * 128 + (r & 7) * 16; }

template <int CIN_, int COUT_, int KS_>
struct Cfg {
    static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_;
    static constexpr int CB = CIN / 16;                    // forward: 16-channel blocks per tap
    static constexpr int NSTEP = KS * KS * CB;
    static constexpr int B_STEP = COUT * 32, B_BYTES = NSTEP * B_STEP;       // forward image: step (tap, ci/16): COUT rows x 16 k
    static constexpr int NW = COUT * CIN * KS * KS;
};
template <typename C>
struct DCfg {                                              // dgrad image: step (tap, co/16): CIN rows x 16 k
    static constexpr int N = C::CIN, CB = C::COUT / 16, NSTEP = C::KS * C::KS * CB;
    static constexpr int B_STEP = N * 32, B_BYTES = NSTEP * B_STEP;
};
using L2 = Cfg<16, 32, 5>;
using L3 = Cfg<32, 64, 4>;
using L4 = Cfg<64, 128, 3>;

// byte offsets of the per-layer operand images inside w_packed: [conv1 | conv2 | conv3 | conv4 | dgrad2 | dgrad3 | dgrad4]
constexpr size_t kPackOff1 = 0, kPackOff2 = 57344;
constexpr size_t kPackOff3 = kPackOff2 + L2::B_BYTES, kPackOff4 = kPackOff3 + L3::B_BYTES;
constexpr size_t kPackD2 = kPackOff4 + L4::B_BYTES;
constexpr size_t kPackD3 = kPackD2 + DCfg<L2>::B_BYTES, kPackD4 = kPackD3 + DCfg<L3>::B_BYTES;
// conv1's Toeplitz image for the swapped-role forward (conv1_fwd4.cu): 36 blocks of 64 rows (co*4 + j) x 16 k, ordered
// [Z][ky=0: ci 3,2,1,0][Z][ky=1: ...]...[Z] -- descending channel with zero blocks between the kernel rows, so that the 128 rows
// [W(ci) ; W(ci-1)] of two consecutive window samples are one contiguous slice for every ci = 0..4
constexpr int kC1V4Blocks = 36, kC1V4Bytes = kC1V4Blocks * 2048;
constexpr size_t kPackC1V4 = kPackD4 + DCfg<L4>::B_BYTES;
// obs_size 12 (BASELINE configs[3]: 3 cameras x 4 frames, channel = 3*frame + camera): conv1 runs as three 4-channel camera
// streams accumulated into one result (conv1_tc.cu), each with its own Toeplitz image [ky][frame]: camera 0 at kPackOff1, 1 and 2 here
constexpr int kC1CamBytes = 57344;
constexpr size_t kPackC1Cam1 = kPackC1V4 + kC1V4Bytes;
constexpr size_t kPackTotal = kPackC1Cam1 + 2 * kC1CamBytes;
__host__ __device__ constexpr size_t c1_cam_off(int cam) { return cam == 0 ? kPackOff1 : kPackC1Cam1 + (size_t)(cam - 1) * kC1CamBytes; }
__host__ __device__ constexpr int c1v4_block(int ky, int ci) { return 1 + 5 * ky + (3 - ci); }
constexpr int kNC1W = 16 * 4 * 7 * 7;                      // conv1 weights (obs_size 4)
constexpr int kNC1W12 = 16 * 12 * 7 * 7;                   // conv1 weights (obs_size 12)

// One source weight W[co][ci][tap] of conv2-4 -> its position in the forward image and in the dgrad image.
template <typename C>
__device__ __forceinline__ void pack_src_elem(float w, __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgr, int i) {
    using D = DCfg<C>;
    constexpr int KK = C::KS * C::KS;
    const int tap = i % KK, ci = (i / KK) % C::CIN, co = i / (KK * C::CIN);
    const __nv_bfloat16 v = __float2bfloat16_rn(w);
    {   // forward: step (tap, ci/16), row co, k = ci%16
        const int sidx = tap * C::CB + (ci >> 4), k = ci & 15;
        fwd[(size_t)sidx * (C::B_STEP / 2) + op_off(co, k >> 3) / 2 + (k & 7)] = v;
    }
    {   // dgrad: row ci, k = co%16. conv2 / conv3 (conv_sw.cu, "Toeplitz in N"): the image is stored FLIPPED and with the taps of a
        // kernel row adjacent -- step ((KS-1-ky) * CB + co/16) * KS + (KS-1-kx) -- so that one MMA reads the KS weight blocks of a
        // window row as N = KS * CIN consecutive rows; conv4 (conv4_sw.cu) keeps step (tap, co/16)
        const int ky = tap / C::KS, kx = tap % C::KS, k = co & 15;
        const int sidx = C::KS >= 4 ? ((C::KS - 1 - ky) * D::CB + (co >> 4)) * C::KS + (C::KS - 1 - kx) : tap * D::CB + (co >> 4);
        dgr[(size_t)sidx * (D::B_STEP / 2) + op_off(ci, k >> 3) / 2 + (k & 7)] = v;
    }
}

// One conv1 weight W[co][ci][ky][kx] (obs_size 4) -> its up to four copies in the Toeplitz image (conv1_tc.cu):
// step (ky,ci): 64 rows n = j*16+co x 16 k, Wt[(j,co)][p] = W[..][p - 3j]; the other entries are structural zeros
// written once by pack_all_kernel.
__device__ __forceinline__ void pack_conv1_elem(float w, __nv_bfloat16* __restrict__ img, __nv_bfloat16* __restrict__ img4, int i) {
    const int kx = i % 7, ky = (i / 7) % 7, ci = (i / 49) & 3, co = i / 196;
    const __nv_bfloat16 v = __float2bfloat16_rn(w);
    const int st = ky * 4 + ci, blk = c1v4_block(ky, ci);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = kx + 3 * j, n = j * 16 + co;
        if (k < 16) {
            img[(size_t)st * 1024 + op_off(n, k >> 3) / 2 + (k & 7)] = v;
            img4[(size_t)blk * 1024 + op_off(co * 4 + j, k >> 3) / 2 + (k & 7)] = v;      // swapped-role image: row = co*4 + j
        }
    }
}

// obs_size 12: W[co][ci = 3*frame + cam][ky][kx] -> image of camera `cam`, step (ky, frame), same Toeplitz rows as above
__device__ __forceinline__ void pack_conv1x12_elem(float w, uint8_t* __restrict__ base, int i) {
    const int kx = i % 7, ky = (i / 7) % 7, ci = (i / 49) % 12, co = i / 588;
    const __nv_bfloat16 v = __float2bfloat16_rn(w);
    __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(base + c1_cam_off(ci % 3));
    const int st = ky * 4 + ci / 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = kx + 3 * j, n = j * 16 + co;
        if (k < 16) img[(size_t)st * 1024 + op_off(n, k >> 3) / 2 + (k & 7)] = v;
    }
}

// Where the conv weights live in the parameter arena (float offsets) + the image base: what an optimiser kernel needs
// to refresh the operand images of the floats it updates. base == nullptr: no refresh (fp32 mode).
struct PackMap {
    uint8_t* base;
    int64_t w1, w2, w3, w4;
    int obs;                 // 4 or 12: which conv1 images exist
};
__host__ inline PackMap pack_map(const bc::Arena& a, void* w_packed, int obs) {
    return PackMap{(uint8_t*)w_packed, a.w[0], a.w[1], a.w[2], a.w[3], obs};
}
// `idx` = arena index of the first of four consecutive floats `v` (tensors are padded to 32 floats: never straddled)
__device__ __forceinline__ void pack_updated4(const PackMap& pm, int64_t idx, const float* v) {
    if (pm.base == nullptr) return;
    if (idx >= pm.w1) {
        const int i = (int)(idx - pm.w1);
        if (pm.obs == 12) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (i + k < kNC1W12) pack_conv1x12_elem(v[k], pm.base, i + k);
            return;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + k < kNC1W) pack_conv1_elem(v[k], (__nv_bfloat16*)pm.base, (__nv_bfloat16*)(pm.base + kPackC1V4), i + k);
    } else if (idx >= pm.w2) {
        const int i = (int)(idx - pm.w2);
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + k < L2::NW) pack_src_elem<L2>(v[k], (__nv_bfloat16*)(pm.base + kPackOff2), (__nv_bfloat16*)(pm.base + kPackD2), i + k);
    } else if (idx >= pm.w3) {
        const int i = (int)(idx - pm.w3);
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + k < L3::NW) pack_src_elem<L3>(v[k], (__nv_bfloat16*)(pm.base + kPackOff3), (__nv_bfloat16*)(pm.base + kPackD3), i + k);
    } else if (idx >= pm.w4) {
        const int i = (int)(idx - pm.w4);
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + k < L4::NW) pack_src_elem<L4>(v[k], (__nv_bfloat16*)(pm.base + kPackOff4), (__nv_bfloat16*)(pm.base + kPackD4), i + k);
    }
}

}  // namespace ctc
