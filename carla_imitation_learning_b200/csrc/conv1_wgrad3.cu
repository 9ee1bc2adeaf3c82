// conv1 weight/bias gradient on tcgen05, third generation (bf16 mode, sliding-window batches, compact gradient inputs).
// Replaces the conv1 part of autograd's convolution_backward for cnn_base[0] of
// /root/reference/src/architectures/nets.py:18-20 (see conv1_tc.cu for the Toeplitz formulation and the second generation,
// which stays the kernel of materialised batches and of the f32 NCHW gradient buffers).
//
// What the profile of the second generation showed (profiles/r2b_*): its four issuer warps -- one per channel role ci, all
// visiting EVERY job (tile row, plane) -- move in lockstep: they wait for the same plane, push their 8 MMAs each, stall on
// the full tensor queue together and then spend the same ~1300 cycles of loop / barrier / commit overhead together, during
// which the tensor pipe idles: 2850 cycles per job for 1555 cycles of MMAs. The builders were waiting for free slots.
// Third generation:
//   * FOUR issuers = (job parity) x (channel pair): issuer (i, pr) owns pair pr = (ci 2pr, 2pr+1) of the jobs k = i mod 2, so a
//     warp visits every second job only and while two issuers are in their overhead phase the other two feed the pipe (the
//     profile of a two-issuer version, profiles/r2h, showed the issuers instruction-bound, the builders waiting for them).
//     Each issuer has a private accumulator (4 x 128 TMEM columns = all of TMEM), added in a fixed order by the epilogue: the
//     result stays bitwise reproducible.
//   * plane P meets dY(P), dY(P-1) (ci = 0, 1) and dY(P-2), dY(P-3) (ci = 2, 3): each pair sits in ADJACENT gradient slots
//     (except across the wrap of the ring), so one MMA of N = 128 does what took two of N = 64 (64.4 instead of 97.2 cycles:
//     tools/mma_bench.py): 16 MMAs per job instead of 32.
//   * a gradient ring of 7 slots (4 live + 3 ahead): the ablation of the first version of this kernel (profiles/r2f) showed the
//     MMAs of a job sitting ON the builder -> issuer -> builder round trip through a 5-slot ring (no overlap at all). The room
//     comes from the A slots: the all-ones blocks that produced the bias gradient as an accumulator row are gone (28 KB instead
//     of 32 KB per plane slot); the builders sum the bias gradient from the values they hold anyway.
//   * the folding epilogue runs on both builder groups (ci 0,1 / ci 2,3): it is a serial tail of the kernel.
//   * the slot bookkeeping of segment edges ("orphan" uses) moved to the builders, who know how many jobs of the segment use
//     a sample and pre-arrive for the missing ones; the issuer loop has no edge logic left beyond valid flags.
//   * A slices at a 2048 B pitch; compact builders (one 16 B + one 8 B load per unit, conv_sw.cu / conv1_tc.cu write them).
#include <stdlib.h>
#include "bc_common.cuh"
#include "tc05.cuh"
#include "trace.cuh"

namespace c1wg3 {


constexpr int NG = 21, TILES_PER_FRAME = 14;
constexpr int NTHREADS = 13 * 32;            // warp 0 loader + TMEM alloc, warps 1-3 and 12 issuers, warps 4-7 / 8-11 builder groups (both fold in the epilogue)
constexpr int ROWB = 336;
constexpr int VIEW = 6 * ROWB;               // one (ky,h) slice as a bulk copy brings it: 126 rows x 16 B
constexpr int DY_BYTES = 64 * 256;           // [8 n-blocks][128 rows][16 B]
constexpr int TP_PIECE_BYTES = 86 * ROWB;
constexpr int TMEM_COLS = 512;
// Shared-memory plan: NA plane slots of 14 slices at pitch VP (the MMA's M groups 14, 15 read what follows the slot: accumulator
// rows 112.. are never used), NDY gradient slots (4 live -- plane P meets dY(P..P-3) -- plus the ones built ahead).
template <int NA_, int NDY_, int VP_>
struct Plan {
    static constexpr int NA = NA_, NDY = NDY_, VPAD = VP_;
    static constexpr int A_SLOT = 14 * VPAD;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_DY = (OFF_A + NA * A_SLOT + 1023) / 1024 * 1024;     // (the over-read of the last slot lands in the gradient ring)
    static constexpr int OFF_BAR = OFF_DY + NDY * DY_BYTES;
    static constexpr int NBAR = 2 * NA + 2 * NDY + 1;
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 64;      // + TMEM address and the 8 'accumulator half written' flags
    static_assert(SMEM_BYTES <= 227 * 1024, "conv1 wgrad v3 shared memory");
};

// One CTA owns a contiguous range of the job list i = ty * (B + 3) + P; a "segment" is the part of it inside one tile row.
struct Seg { int ty, Pa, Pb, sa, sb; };
struct SegIter {
    int i, hi, NJ, B;
    __device__ SegIter(int B_) : B(B_) {
        NJ = B_ + 3;
        const long long T = (long long)NJ * TILES_PER_FRAME;
        i = (int)(T * blockIdx.x / gridDim.x);
        hi = (int)(T * (blockIdx.x + 1) / gridDim.x);
    }
    __device__ bool next(Seg& s) {
        if (i >= hi) return false;
        s.ty = i / NJ; s.Pa = i - s.ty * NJ;
        const int n = min(hi - i, NJ - s.Pa);
        s.Pb = s.Pa + n - 1;
        s.sa = max(0, s.Pa - 3); s.sb = min(B - 1, s.Pb);
        i += n;
        return true;
    }
};

template <typename PL, bool CAM3>
__global__ void __launch_bounds__(NTHREADS, 1)
conv1_wgrad3_kernel(const __nv_bfloat16* __restrict__ x, int64_t sc, const uint4* __restrict__ g8, const uint2* __restrict__ a8,
                    float* __restrict__ part, int64_t seg_len, int64_t w_off, int64_t b_off, int nparts, int B, int* err, int ablate,
                    int cam) {
    // the 4-channel network is CAM3 = false with compile-time (cin, cstep, coff) = (4, 1, 0) -- as runtime values they cost the
    // builders registers (measured: 29.8 -> 32.3 us). obs_size 12 (3 cameras x 4 frames): three launches over the cameras' own
    // sliding windows (x + cam planes, plane stride 3), launch `coff` = cam fills the network's channels 3*ci + cam of every
    // partial slot; the bias gradient and the zeroing of the slots no CTA owns belong to launch 0.
    constexpr int cin = CAM3 ? 12 : 4, cstep = CAM3 ? 3 : 1;
    const int coff = CAM3 ? cam : 0;
    constexpr int NA = PL::NA, NDY = PL::NDY, VPAD = PL::VPAD, A_SLOT = PL::A_SLOT, OFF_A = PL::OFF_A, OFF_DY = PL::OFF_DY,
                  OFF_BAR = PL::OFF_BAR, NBAR = PL::NBAR;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a_full = bars;                  // [NA]   loader (tx)
    uint64_t* a_empty = bars + NA;            // [NA]   the two owning issuers' commits
    uint64_t* dy_full = bars + 2 * NA;        // [NDY]  4 builder warps
    uint64_t* dy_empty = bars + 2 * NA + NDY; // [NDY]  4 = one per job that uses the sample (+ the builders' pre-arrivals at segment edges)
    uint64_t* done = bars + 2 * NA + 2 * NDY; //        4 issuers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    volatile uint32_t* written = reinterpret_cast<volatile uint32_t*>(tmem_slot) + 1;     // [job parity][pair][half]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NA; ++i) { tc05::mbar_init(a_full + i, 1); tc05::mbar_init(a_empty + i, 2); }
        for (int i = 0; i < NDY; ++i) { tc05::mbar_init(dy_full + i, 4); tc05::mbar_init(dy_empty + i, 4); }
        tc05::mbar_init(done, 4);
        tc05::mbar_fence_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, TMEM_COLS);
    // A slots start as zeros (the 32 B of padding behind every slice = K rows 126,127: never a NaN pattern); the dY ring too:
    // rows 0..125 of a slot are rewritten for every sample tile, rows 126,127 stay zero
    for (int i = threadIdx.x; i < (OFF_DY + NDY * DY_BYTES) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc05::fence_async_smem();
    tc05::tc_fence_before();
    __syncthreads();
    tc05::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TRACE_T0
    TRACE_DECL
    TRACE(0, 0);
    tc05::pdl_trigger();
    tc05::pdl_wait();
    TRACE(10, 0);

    if (warp == 0) {
        // ------------------------------------------------------------------ loader: 14 bulk copies per job, one per lane
        SegIter it(B);
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
                const uint8_t* src = reinterpret_cast<const uint8_t*>(x + (int64_t)P * sc) + (size_t)(6 * sg.ty) * ROWB;
                uint8_t* dst = smem + OFF_A + slot * A_SLOT;
                if (ablate & 2) {                // timing experiment (-DBC_ABLATE builds only): no plane loads
                    if (lane == 0) tc05::mbar_arrive(a_full + slot);
                    continue;
                }
                TRACE(1, k);
                if (lane == 0) tc05::mbar_expect_tx(a_full + slot, 14 * VIEW);
                __syncwarp();
                if (lane < 14) tc05::bulk_g2s(dst + dst_off, src + src_off, VIEW, a_full + slot);
            }
        }
    } else if (warp <= 3 || warp == 12) {
        // ------------------------------------------------------------------ issuer (me, pr): pair pr of every second job
        const uint32_t q = warp == 12 ? 3u : (uint32_t)(warp - 1);
        const uint32_t me = q & 1u, pr = q >> 1;
        constexpr uint32_t id64 = tc05::instr_desc(tc05::FMT_BF16, 128, 64, 1, 1);       // MN-major A and B
        constexpr uint32_t id128 = tc05::instr_desc(tc05::FMT_BF16, 128, 128, 1, 1);
        const uint64_t ad0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_A), 128, VPAD, tc05::SW_NONE);
        const uint64_t bd0 = tc05::smem_desc(tc05::smem_u32(smem + OFF_DY), 128, 2048, tc05::SW_NONE);
        const uint32_t d = tmem_base + (me * 2 + pr) * 128;     // columns 0..63 = the odd ci of the pair (older sample), 64..127 = the even ci
        SegIter it(B);
        Seg sg;
        uint32_t k = 0, kb0 = 0;               // job counter, build counter at the start of the segment
        bool f_hi = true, f_lo = true;         // that accumulator half has not been written yet (hi = even ci)
        bool ok = true;
        while (ok && it.next(sg)) {
            for (int P = sg.Pa; ok && P <= sg.Pb; ++P, ++k) {
                if ((k & 1u) != me) continue;
                const uint32_t slot = k % NA, ph = (k / NA) & 1;
                ok = tc05::mbar_wait(a_full + slot, ph, err);
                const int s_hi = P - 2 * (int)pr, s_lo = s_hi - 1;          // even ci = newer sample, odd ci = older
                const bool v_hi = s_hi >= sg.sa && s_hi <= sg.sb, v_lo = s_lo >= sg.sa && s_lo <= sg.sb;
                const uint32_t kb_hi = kb0 + (uint32_t)(s_hi - sg.sa), kb_lo = kb_hi - 1;
                const uint32_t d_hi = kb_hi % NDY, d_lo = kb_lo % NDY;
                if (v_lo) ok = ok && tc05::mbar_wait(dy_full + d_lo, (kb_lo / NDY) & 1, err);
                if (v_hi) ok = ok && tc05::mbar_wait(dy_full + d_hi, (kb_hi / NDY) & 1, err);
                tc05::tc_fence_after();
                TRACE(3, k);
                if (ok && tc05::elect_one()) {
                    const uint64_t a_st = ad0 + (uint64_t)((slot * A_SLOT) >> 4);
                    const uint64_t b_lo = bd0 + (uint64_t)((d_lo * DY_BYTES) >> 4);      // the older sample's slot; the newer one is the
                    const uint64_t b_hi = bd0 + (uint64_t)((d_hi * DY_BYTES) >> 4);      // next slot unless the ring wraps between them
                    if (!(ablate & 1)) {                                                 // ablate bit 0: no MMAs (commits only)
                        if (v_hi && v_lo && f_hi == f_lo && d_hi == d_lo + 1) {
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                tc05::mma_bf16(d, a_st + (uint64_t)(u * 16), b_lo + (uint64_t)(u * 16), id128, (f_lo && u == 0) ? 0u : 1u);
                        } else {
                            if (v_lo) {
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    tc05::mma_bf16(d, a_st + (uint64_t)(u * 16), b_lo + (uint64_t)(u * 16), id64, (f_lo && u == 0) ? 0u : 1u);
                            }
                            if (v_hi) {
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    tc05::mma_bf16(d + 64, a_st + (uint64_t)(u * 16), b_hi + (uint64_t)(u * 16), id64, (f_hi && u == 0) ? 0u : 1u);
                            }
                        }
                    }
                    tc05::mma_commit(a_empty + slot);
                    if (v_lo) tc05::mma_commit(dy_empty + d_lo);
                    if (v_hi) tc05::mma_commit(dy_empty + d_hi);
                }
                __syncwarp();
                TRACE(4, k);
                f_hi = f_hi && !v_hi;
                f_lo = f_lo && !v_lo;
            }
            kb0 += (uint32_t)(sg.sb - sg.sa + 1);
        }
        // tell the epilogue which accumulator halves were ever written (an unwritten half holds garbage)
        if (lane == 0) { written[me * 4 + 2 * pr] = f_lo ? 0u : 1u; written[me * 4 + 2 * pr + 1] = f_hi ? 0u : 1u; }
        __threadfence_block();
        __syncwarp();
        if (tc05::elect_one()) tc05::mma_commit(done);
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ two dY builder groups (alternate samples), then the folding epilogue
        const int grp = warp >= 8 ? 1 : 0;
        const int ew = warp & 3;
        const int te = threadIdx.x - (grp ? 256 : 128);
        // One unit = (pooled pixel of the tile, 8 channels): its 3x3 conv positions x 8 channels are nine full 16 B stores
        // (zeros except each channel's routed position), so a sample tile REWRITES rows 0..125 of its slot completely.
        const int upx = te % 28, ucg = (te / 28) & 1, upyl = te / 56;       // units 0..111 of the 128 threads
        const bool unit = te < 112;
        struct Build {
            SegIter it; Seg sg; int smp; uint32_t kb; bool valid;
            __device__ Build(int B_) : it(B_), smp(0), kb(0) { valid = it.next(sg); if (valid) smp = sg.sa; }
            __device__ void step() {                       // next build of the CTA
                ++kb;
                if (++smp > sg.sb) { valid = it.next(sg); if (valid) smp = sg.sa; }
            }
            __device__ void step_group() { step(); if (valid) step(); }
        };
        uint4 ga, gb; uint2 aa, ab;                                        // two register sets: loads run two group-builds ahead of the stores
        ga = gb = make_uint4(0, 0, 0, 0); aa = ab = make_uint2(0xffffffffu, 0xffffffffu);
        auto fetch = [&](uint4& g, uint2& a, const Build& bd) {
            if (!unit || !bd.valid) return;
            const size_t idx = ((size_t)bd.smp * 2 + ucg) * 784 + (2 * bd.sg.ty + upyl) * 28 + upx;
            g = __ldg(g8 + idx);
            a = __ldg(a8 + idx);
        };
        float bsum[8];                                                     // bias gradient: sum of this thread's 8 channels over its builds
#pragma unroll
        for (int q = 0; q < 8; ++q) bsum[q] = 0.f;
        auto store = [&](const uint4& g, const uint2& a, uint8_t* dy, bool count_bias) {
            if (!unit) return;
            if (count_bias) {      // every channel's value lands at exactly one of the 9 positions: the bias gradient is their plain sum
                const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    bsum[2 * q] += __uint_as_float(gw[q] << 16);
                    bsum[2 * q + 1] += __uint_as_float(gw[q] & 0xffff0000u);
                }
            }
#pragma unroll
            for (int p = 0; p < 9; ++p) {
                // per-byte equality masks of the 8 routing codes against p, widened to the bf16 halves they select
                const uint32_t m_lo = __vcmpeq4(a.x, 0x01010101u * (uint32_t)p), m_hi = __vcmpeq4(a.y, 0x01010101u * (uint32_t)p);
                const uint4 w = make_uint4(g.x & __byte_perm(m_lo, 0, 0x1100), g.y & __byte_perm(m_lo, 0, 0x3322),
                                           g.z & __byte_perm(m_hi, 0, 0x1100), g.w & __byte_perm(m_hi, 0, 0x3322));
                const int oyl = 3 * upyl + p / 3, ox = 3 * upx + p % 3;
                // row r = (oyl, ox / 4), column block = (ox % 4) * 2 + channel half: [n/8][128 rows][16 B]
                const int off = ((ox & 3) * 2 + ucg) * 2048 + (oyl * NG + (ox >> 2)) * 16;
                *reinterpret_cast<uint4*>(dy + off) = w;
            }
        };
        Build cur(B);
        if (grp && cur.valid) cur.step();                  // group 1 starts at the CTA's second build
        Build ahead = cur;
        fetch(ga, aa, ahead);
        if (ahead.valid) ahead.step_group();
        fetch(gb, ab, ahead);
        if (ahead.valid) ahead.step_group();               // `ahead` = two group-builds past `cur`
        bool ok = true;
#pragma unroll 1
        for (uint32_t n = 0; ok && cur.valid; ++n) {
            const uint32_t slot = cur.kb % NDY;
            TRACE(5, cur.kb);
            ok = tc05::mbar_wait(dy_empty + slot, ((cur.kb / NDY) & 1) ^ 1, err);
            if (!ok) break;
            TRACE(6, cur.kb);
            uint8_t* dy = smem + OFF_DY + slot * DY_BYTES;
            // a (sample, tile) is counted once for the bias gradient: by the CTA that owns the job in which it is channel role 0
            const bool cb = cur.smp >= cur.sg.Pa && cur.smp <= cur.sg.Pb;
            if (!(ablate & 4)) { if (n & 1) store(gb, ab, dy, cb); else store(ga, aa, dy, cb); }     // ablate bit 2: no gradient stores
            tc05::fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tc05::mbar_arrive(dy_full + slot);
                if (ew == 0) {
                    // jobs of this segment that use the sample: P = smp + ci inside [Pa, Pb]; the missing ones (segment edges) never
                    // arrive on dy_empty, so their arrivals are made here, for the phase this build has just opened
                    const int uses = min(cur.sg.Pb, cur.smp + 3) - max(cur.sg.Pa, cur.smp) + 1;
                    for (int i = uses; i < 4; ++i) tc05::mbar_arrive(dy_empty + slot);
                }
            }
            TRACE(7, cur.kb);
            if (!(ablate & 8)) { if (n & 1) fetch(gb, ab, ahead); else fetch(ga, aa, ahead); }   // the set just stored is free: load two group-builds ahead (ablate bit 3: no gradient loads)
            cur.step_group();
            if (ahead.valid) ahead.step_group();
        }
        // ---- epilogue (both groups; group g takes ci = 2g, 2g+1): add the two issuers' accumulators (fixed order), fold the
        // Toeplitz rows back to 7 taps and write this CTA's partial in arena order; then the bias gradient from the builders' sums
        float* dst = part + (size_t)blockIdx.x * seg_len;
        const int nw = 16 * cin * 49;
        if (coff == 0 && grp == 0 && (int)blockIdx.x + (int)gridDim.x < nparts) {           // slots this launch does not own must read as zero
            for (int s2 = blockIdx.x + gridDim.x; s2 < nparts; s2 += gridDim.x)
                for (int i = te; i < nw + 16; i += 128) part[(size_t)s2 * seg_len + (i < nw ? w_off + i : b_off + i - nw)] = 0.f;
        }
        TRACE(8, 0);
        const bool fin = ok && tc05::mbar_wait(done, 0, err);
        TRACE(9, 0);
        if (fin && !(ablate & 16)) {              // ablate bit 4: no epilogue
            tc05::tc_fence_after();
            const int ky = 2 * ew + (lane >> 4), p = lane & 15;     // accumulator row m = ky*16 + p
            const bool wrow = ky < 7 && p < 7;                      // this lane writes tap kx = p
#pragma unroll 1
            for (int ci = 2 * grp; ci < 2 * grp + 2; ++ci) {
                float v[64];
#pragma unroll
                for (int c = 0; c < 64; ++c) v[c] = 0.f;
                const int pair = ci >> 1, half = (ci & 1) ? 0 : 1;
#pragma unroll 1
                for (int is = 0; is < 2; ++is) {
                    if (!written[is * 4 + 2 * pair + half]) continue;
                    float t[64];
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 16)
                        tc05::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + is * 256 + pair * 128 + half * 64 + c0, t + c0);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 64; ++c) v[c] += t[c];
                }
#pragma unroll
                for (int co = 0; co < 16; ++co) {
                    float a = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // dW[.., kx] += dWt[p = 3j + kx][(j, co)]: fetch column j*16+co from the lane holding row 3j+kx
                        const int srcl = (lane & 16) + ((3 * j + p) & 15);
                        a += __shfl_sync(0xffffffffu, v[j * 16 + co], srcl);
                    }
                    if (wrow) dst[w_off + ((size_t)(co * cin + ci * cstep + coff) * 7 + ky) * 7 + p] = a;
                }
            }
        }
        // bias gradient: the threads' sums through the (now idle) gradient ring, added in a fixed order
        float* scratch = reinterpret_cast<float*>(smem + OFF_DY);
        if (fin && unit) {
#pragma unroll
            for (int q = 0; q < 8; ++q) scratch[(grp * 112 + te) * 8 + q] = bsum[q];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (fin && coff == 0 && grp == 0) {
            // 16 channels x 112 partial sums each (2 groups x 56 units of the channel's group). Eight threads per channel take 14 values
            // each in a fixed order and meet in a fixed xor tree -- the CTA timeline (profiles/r2U) showed the earlier version, one thread
            // per channel walking all 224 scratch rows, as a 7 us serial tail of the kernel
            const int ch = te >> 3, part = te & 7, cg = ch >> 3, q = ch & 7;
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                const int v = part + 8 * i, g2 = v / 56, w = v % 56;
                const int u = (w / 28) * 56 + 28 * cg + w % 28;
                a += scratch[(g2 * 112 + u) * 8 + q];
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            if (part == 0) dst[b_off + ch] = a;
        }
    }
    TRACE(11, 0);
    TRACE_END;
    tc05::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace c1wg3

BC_TRACE_EXPORT(bc_debug_c1wg3_trace)      // debug builds only: not part of the ABI

// sliding-window batch + compact gradient inputs (bc_ctx.conv_mode bit 16); everything else: conv1_tc.cu's kernel
int bc_conv1_wgrad3_launch(const bc_ctx* c, const bc::Arena& ar, const bc::Partials& pl, int grid, void* stream) {
    using P47 = c1wg3::Plan<4, 7, 2048>;
    using P55 = c1wg3::Plan<5, 5, 2016>;
    auto k47 = c1wg3::conv1_wgrad3_kernel<P47, false>;
    auto k55 = c1wg3::conv1_wgrad3_kernel<P55, false>;
    auto k47c = c1wg3::conv1_wgrad3_kernel<P47, true>;      // obs_size 12: one launch per camera
    static bc::PerDeviceOnce once_; bool& configured = once_();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k47, cudaFuncAttributeMaxDynamicSharedMemorySize, P47::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k55, cudaFuncAttributeMaxDynamicSharedMemorySize, P55::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k47c, cudaFuncAttributeMaxDynamicSharedMemorySize, P47::SMEM_BYTES);
        if (e != cudaSuccess) return bc::fail(BC_ERR_DEVICE, "conv1 wgrad v3: smem opt-in failed: %s", cudaGetErrorString(e));
        configured = true;
    }
#ifdef BC_ABLATE   // timing experiments only (knowingly wrong results): compiled out of the shipped library
    static const int ablate = getenv("BC_C1WG_ABLATE") ? atoi(getenv("BC_C1WG_ABLATE")) : 0;
#else
    constexpr int ablate = 0;
#endif
    static const bool plan55 = getenv("BC_C1WG_PLAN") && atoi(getenv("BC_C1WG_PLAN")) == 55;    // measurement switch (same results)
    const int ncam = c->obs_size == 12 ? 3 : 1;
    for (int cam = 0; cam < ncam; ++cam) {
        bc::launch_pdl(ncam == 3 ? k47c : plan55 ? k55 : k47, dim3(grid), dim3(c1wg3::NTHREADS), (plan55 && ncam == 1) ? P55::SMEM_BYTES : P47::SMEM_BYTES, (cudaStream_t)stream,
            (const __nv_bfloat16*)c->x_tp + (int64_t)cam * c->x_tp_stride_c, (int64_t)ncam * c->x_tp_stride_c, (const uint4*)c->gact0_p8, (const uint2*)c->amax0_p8,
            c->partials + pl.off[4], ar.seg_len[4], ar.w[0] - ar.seg_off[4], ar.b[0] - ar.seg_off[4], grid, c->batch, c->err_flag, ablate,
            cam);
        BC_CUDA_LAUNCH_CHECK("conv1_wgrad3_kernel");
    }
    return BC_OK;
}
