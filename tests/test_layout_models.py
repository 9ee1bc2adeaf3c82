"""CPU models of the operand layouts the tcgen05 kernels rely on (include/bc_b200.h BC_BF16_TP, csrc/conv1_tc.cu,
csrc/conv_sw.cu): the descriptor arithmetic written out in numpy must reproduce the convolution's operands exactly.
Pure index identities -- no GPU, no oracle numerics -- so a layout change that breaks a kernel's assumption fails here first."""
import numpy as np
import pytest

ROWB_ELEMS = 21 * 8            # one (q) row of a TP piece: 21 groups x 8 pixels


def tp_plane(img):
    """(256,256) -> (3,2,86,21,8): TP[c][h][q][g][i] = img[3q+c][12g+8h+i], zero rows beyond 255."""
    pad = np.zeros((258, 256), img.dtype)
    pad[:256] = img
    R = 3 * np.arange(86)[None, :] + np.arange(3)[:, None]
    px = 12 * np.arange(21)[None, :, None] + 8 * np.arange(2)[:, None, None] + np.arange(8)[None, None, :]
    return pad[R[:, None, :, None, None], px[None, :, None, :, :]]


@pytest.mark.parametrize("ty", [0, 5, 13])
def test_conv1_a_operand_is_a_strided_view_of_the_tp_tile(ty):
    """conv1_tp_kernel: the slot holds pieces (c,h) = rows q0..q0+nq(c)-1; kernel row ky = 3d + c reads the 126x16 slice
    starting d rows in, 16 B row pitch (SBO 128 B per 8 rows), K halves one piece apart (LBO)."""
    rng = np.random.default_rng(ty)
    img = rng.integers(1, 1 << 15, size=(256, 256)).astype(np.int32)
    tp = tp_plane(img)
    nq = (8, 7, 7)
    slot, off = [], {}
    for c in range(3):
        for h in range(2):
            off[c, h] = sum(len(p) for p in slot)
            slot.append(tp[c, h, 6 * ty: 6 * ty + nq[c]].reshape(-1))
    slot = np.concatenate(slot + [np.zeros(64, np.int32)])          # + the over-read pad
    for ky in range(7):
        c, d = ky % 3, ky // 3
        start, lbo = off[c, 0] + d * ROWB_ELEMS, off[c, 1] - off[c, 0]
        r, k = np.arange(126)[:, None], np.arange(16)[None, :]
        A = slot[start + (k // 8) * lbo + r * 8 + k % 8]            # UMMA K-major, no swizzle: row pitch 16 B = 8 elements
        want = img[3 * (6 * ty + r // 21) + ky, 12 * (r % 21) + k]
        assert np.array_equal(A, want), ky


def test_conv1_wgrad_slices_have_a_uniform_stride():
    """conv1_wgrad_tp_kernel: the 14 (ky,h) slices = rows d..d+5 of TP piece (c,h) are contiguous 2016 B runs in HBM, so
    14 bulk copies can lay them at a uniform stride (the M-direction SBO of the MN-major operand)."""
    img = np.arange(256 * 256, dtype=np.int32).reshape(256, 256)
    tp = tp_plane(img)
    ty = 7
    for ky in range(7):
        c, d = ky % 3, ky // 3
        for h in range(2):
            piece = tp[c, h].reshape(-1)                              # (q, g, i) contiguous in HBM
            run = piece[(6 * ty + d) * ROWB_ELEMS: (6 * ty + d + 6) * ROWB_ELEMS]
            r, i = np.arange(126)[:, None], np.arange(8)[None, :]
            assert np.array_equal(run.reshape(126, 8), img[3 * (6 * ty + r // 21) + ky, 12 * (r % 21) + 8 * h + i])


def p8(x):
    """(C,H,W) -> (C/8, H*W, 8): the shifted-window kernels' activation layout."""
    C, H, W = x.shape
    return x.reshape(C // 8, 8, H * W).transpose(0, 2, 1)


def test_shifted_window_forward_operand():
    """conv_sw.cu forward: row m = oy*W_in + ox; tap (ky,kx), channel block cb = the same image shifted by ky*W_in + kx
    pixels, planes 2cb and 2cb+1 (LBO = one plane)."""
    rng = np.random.default_rng(0)
    C, H, K = 16, 28, 5
    x = rng.integers(1, 1 << 15, size=(C, H, H)).astype(np.int32)
    img = np.concatenate([p8(x).reshape(-1), np.zeros(4096, np.int32)])
    plane = H * H * 8
    m = np.arange(24 * H)[:, None]                                    # all conv rows, pitch W_in
    k = np.arange(16)[None, :]
    oy, ox = m // H, m % H
    for ky, kx, cb in ((0, 0, 0), (4, 4, 0), (2, 3, 0)):
        A = img[(2 * cb + k // 8) * plane + (m + ky * H + kx) * 8 + k % 8]
        valid = (ox < H - K + 1)[:, 0]
        want = x[16 * cb + k, oy + ky, np.minimum(ox + kx, H - 1)]
        assert np.array_equal(A[valid], want[valid])


def test_shifted_window_dgrad_is_a_correlation_of_the_padded_gradient():
    """dX[iy][ix] = sum dYp[iy+ky'][ix+kx'] * W[K-1-ky'][K-1-kx'] with dYp = dY zero-padded by K-1 (size H_in + K - 1)."""
    rng = np.random.default_rng(1)
    H, K = 12, 4
    Ho = H - K + 1
    dy = rng.standard_normal((Ho, Ho))
    w = rng.standard_normal((K, K))
    dx = np.zeros((H, H))
    for oy in range(Ho):
        for ox in range(Ho):
            dx[oy:oy + K, ox:ox + K] += dy[oy, ox] * w              # transpose of the forward correlation
    WP = H + K - 1
    dyp = np.zeros((WP, WP))
    dyp[K - 1:K - 1 + Ho, K - 1:K - 1 + Ho] = dy
    got = np.zeros((H, H))
    for ky in range(K):
        for kx in range(K):
            got += dyp[ky:ky + H, kx:kx + H] * w[K - 1 - ky, K - 1 - kx]
    assert np.allclose(got, dx, atol=1e-12)


def test_shifted_window_wgrad_cores_are_pixel_shifts():
    """conv_sw.cu wgrad: MN-major A with M-cores 16 B apart: row (kx', ci8), K index = pixel m reads
    plane[m + ky*W_in + kx'][ci8]; with dY zero where ox >= W_out the product is the weight gradient."""
    rng = np.random.default_rng(2)
    C, H, K = 8, 12, 4
    Ho = 8                                                            # pooled conv region
    x = rng.standard_normal((C, H, H))
    dy = np.zeros((Ho, H))
    dy[:, :Ho] = rng.standard_normal((Ho, Ho))                       # linear pitch W_in, zero beyond the conv width
    plane = np.concatenate([p8(x)[0].reshape(-1), np.zeros(64 * 8)])
    m = np.arange(Ho * H)
    for ky in range(K):
        A = np.stack([plane[(m + ky * H + kx)[:, None] * 8 + np.arange(8)[None, :]] for kx in range(K)])   # (kx', m, ci8)
        got = np.einsum("kmc,m->kc", A, dy.reshape(-1))
        want = np.array([[np.sum(dy[:, :Ho] * x[ci, ky:ky + Ho, kx:kx + Ho]) for ci in range(8)] for kx in range(K)])
        assert np.allclose(got, want, atol=1e-10)


@pytest.mark.parametrize("S", [1, 2, 3, 4])
def test_conv1_merged_n_mma_gives_every_sample_its_own_convolution(S):
    """conv1_tp_kernel issuer: plane j of a super-tile is channel ci = j - s of every sample s in [max(0, j-3), min(j, S-1)].
    With the weight image ordered [ky][ci][64 rows] and the accumulators of the super-tile in column block 3 - s, ONE
    instruction per (plane, ky) covers the block of samples: B rows start at (ky, ci_lo = j - s_hi), D columns at block
    3 - s_hi, N = 64 * n. The first instruction of a sample (ci = 0, ky = 0) overwrites, split off the block.
    Model: integer operands, the instruction sequence exactly as issued; every sample must end with
    sum_{ci,ky} A(plane s + ci, ky) @ Wt[ci, ky].T and nothing else (stale accumulator contents must vanish)."""
    rng = np.random.default_rng(S)
    M, K, N = 8, 16, 64                                      # rows shrunk (the row dimension plays no role in the mapping)
    A = rng.integers(-3, 4, size=(S + 3, 7, M, K))           # [plane][ky] operand slices
    Wt = rng.integers(-3, 4, size=(4, 7, N, K))              # [ci][ky] Toeplitz weight blocks
    Bimg = Wt.transpose(1, 0, 2, 3).reshape(7 * 4 * N, K)    # packed image: step (ky, ci) -> 64 consecutive rows
    D = rng.integers(-99, 100, size=(M, 256))                # one accumulator set, stale contents
    n_instr = 0
    for j in range(S + 3):
        s_lo, s_hi = max(0, j - 3), min(j, S - 1)
        n, ci_lo, fresh = s_hi - s_lo + 1, j - s_hi, j <= S - 1
        d0 = (3 - s_hi) * 64
        for ky in range(7):
            b0 = (ky * 4 + ci_lo) * N
            if ky == 0 and fresh:
                assert ci_lo == 0                           # the sample that starts here is the first block of the run
                D[:, d0:d0 + 64] = A[j, ky] @ Bimg[b0:b0 + 64].T
                n_instr += 1
                if n > 1:
                    D[:, d0 + 64:d0 + 64 * n] += A[j, ky] @ Bimg[b0 + 64:b0 + 64 * n].T
                    n_instr += 1
            else:
                D[:, d0:d0 + 64 * n] += A[j, ky] @ Bimg[b0:b0 + 64 * n].T
                n_instr += 1
    for s in range(S):
        want = sum(A[s + ci, ky] @ Wt[ci, ky].T for ci in range(4) for ky in range(7))
        assert np.array_equal(D[:, (3 - s) * 64:(4 - s) * 64], want), s
    assert n_instr == 7 * (S + 3) + S - 1                     # vs 28 * S one-sample instructions


def test_stacked_camera_conv_is_the_sum_of_three_four_frame_camera_streams():
    """csrc/conv1_tc.cu (obs_size 12): with three cameras interleaved frame by frame, channel 3f + cam of sample s is plane
    3(s + f) + cam, so camera `cam` alone is a 4-frame sliding window over planes cam, cam + 3, ... (both strides 3 planes) with
    the weights W[:, 3f + cam], and the 12-channel convolution is the sum of the three 4-channel ones -- integer arithmetic, exact."""
    rng = np.random.default_rng(12)
    B, Hh, K, S = 3, 25, 7, 3
    planes = rng.integers(-8, 9, size=(3 * (B - 1) + 12, Hh, Hh)).astype(np.int64)
    W = rng.integers(-4, 5, size=(16, 12, K, K)).astype(np.int64)
    ho = (Hh - K) // S + 1

    def conv(x, w):            # x (C,H,H), w (16,C,K,K) -> (16,ho,ho)
        out = np.zeros((16, ho, ho), np.int64)
        for oy in range(ho):
            for ox in range(ho):
                out[:, oy, ox] = np.tensordot(w, x[:, S * oy:S * oy + K, S * ox:S * ox + K], axes=([1, 2, 3], [0, 1, 2]))
        return out
    for s in range(B):
        full = conv(planes[3 * s:3 * s + 12], W)
        acc = np.zeros_like(full)
        for cam in range(3):
            stream = planes[cam::3]                                  # the camera's own plane sequence
            acc += conv(stream[s:s + 4], W[:, cam::3])               # its 4-frame window of sample s, channels 3f + cam
        assert np.array_equal(acc, full)
        # and the weight gradient splits the same way: dW[:, 3f + cam] is the 4-channel wgrad of camera cam's stream
    dY = rng.integers(-3, 4, size=(B, 16, ho, ho)).astype(np.int64)

    def wgrad(xs, dys):        # list of (C,H,H), (16,ho,ho) -> (16,C,K,K)
        g = np.zeros((16, xs[0].shape[0], K, K), np.int64)
        for x, dy in zip(xs, dys):
            for oy in range(ho):
                for ox in range(ho):
                    g += dy[:, oy, ox][:, None, None, None] * x[None, :, S * oy:S * oy + K, S * ox:S * ox + K]
        return g
    full = wgrad([planes[3 * s:3 * s + 12] for s in range(B)], dY)
    for cam in range(3):
        stream = planes[cam::3]
        assert np.array_equal(wgrad([stream[s:s + 4] for s in range(B)], dY), full[:, cam::3])


def test_swapped_role_conv1_blocks_and_quad_exchange():
    """csrc/conv1_fwd4.cu (BC_C1FW_GEN=4). (a) Weight image order [Z][ky: ci 3,2,1,0][Z]...: for every plane offset d = 0..4 of a
    sample pair the 128 A rows [W(ci = d) ; W(ci = d - 1)] are two consecutive blocks, zero blocks standing in for ci = 4 and -1.
    (b) Quad exchange: lane j of a channel's quad holds conv columns 4g + j; after rotating its three group values by
    r0 = (3 * (3j % 4)) // 4 and shuffling from lane (3L + t) % 4, lane L holds exactly columns 3L .. 3L+2 of the 12-column period."""
    block = lambda ky, ci: 1 + 5 * ky + (3 - ci)
    for ky in range(7):
        for d in range(5):
            hi, lo = block(ky, d), block(ky, d) + 1
            assert hi == (5 * ky if d == 4 else block(ky, d))        # ci = 4 -> the zero block in front of the kernel row
            assert lo == (5 * (ky + 1) if d == 0 else block(ky, d - 1))   # ci = -1 -> the zero block behind it
            assert all(b % 5 != 0 for b, c in ((hi, d), (lo, d - 1)) if 0 <= c <= 3)
    cols = np.arange(12)                                             # one period: value = its conv column
    V = {j: [cols[4 * g + j] for g in range(3)] for j in range(4)}   # lane j: group registers g = 0..2
    for L in range(4):
        got = []
        for t in range(3):
            src = (3 * L + t) % 4                                    # the lane asked at step t
            r0 = (3 * ((3 * src) % 4)) // 4
            got.append(V[src][(r0 + t) % 3])                         # what that lane sends at step t
        assert got == [3 * L, 3 * L + 1, 3 * L + 2]


@pytest.mark.parametrize("hin,ks,wp", [(28, 5, 32), (12, 4, 16)])
def test_dgrad_toeplitz_in_n_fold(hin, ks, wp):
    """csrc/conv_sw.cu dgrad, "Toeplitz in N": per window row ky' ONE product of the padded gradient at kx' = 0 with the KS flipped
    weight blocks side by side; block j of pixel m is a contribution to pixel m - j, so dX[m] = sum_j D[m + j][block j], and m + j
    never leaves the image row of a valid pixel (pitch wp >= hin + ks - 1, 128 % wp == 0: tiles and warps are whole rows)."""
    rng = np.random.default_rng(hin)
    cin, cout, ho = 3, 2, hin - ks + 1
    dy = rng.integers(-3, 4, size=(cout, ho, ho)).astype(np.int64)
    W = rng.integers(-3, 4, size=(cout, cin, ks, ks)).astype(np.int64)
    # direct definition: dX[ci][iy][ix] = sum dY[co][iy-ky][ix-kx] W[co][ci][ky][kx]
    want = np.zeros((cin, hin, hin), np.int64)
    for ky in range(ks):
        for kx in range(ks):
            want[:, ky:ky + ho, kx:kx + ho] += np.einsum("oyx,oc->cyx", dy, W[:, :, ky, kx])
    # kernel: padded gradient (ks-1 zeros around), linear pixel index with pitch wp
    pad = np.zeros((cout, wp + ks, wp), np.int64)
    pad[:, ks - 1:ks - 1 + ho, ks - 1:ks - 1 + ho] = dy
    lin = pad.reshape(cout, -1)
    nrows = hin * wp
    D = np.zeros((nrows + ks, ks, cin), np.int64)              # accumulator rows x column blocks
    for kyp in range(ks):
        A = lin[:, kyp * wp: kyp * wp + nrows + ks].T          # row m = gradient pixel m + ky' wp (window column 0)
        for j in range(ks):                                     # block j = flipped tap (ks-1-ky', ks-1-j)
            D[:, j] += A @ W[:, :, ks - 1 - kyp, ks - 1 - j]
    got = np.zeros_like(want)
    for iy in range(hin):
        for ix in range(hin):
            m = iy * wp + ix
            assert (m + ks - 1) // wp == iy and (m % 32) + ks - 1 < 32          # the fold stays in the row and in the warp
            got[:, iy, ix] = sum(D[m + j, j] for j in range(ks))
    assert np.array_equal(got, want)
    assert 128 % wp == 0 and wp >= hin + ks - 1


def test_policy_tail_thread_mapping_reproduces_conv3_conv4_head():
    """csrc/policy_tail.cu written out thread by thread in numpy: the shared-memory plan (padded channel pitches W3P / W4P / PTP /
    A3P), the (2 channels x pooled row x 2 input channels) register tile of conv3, the partial reduction split over (px, dy, h, g)
    lanes, the pull all-gather across the 8 CTAs of a cluster, conv4's (channel, 4 input channels) threads with the 16-lane fold, and
    the lane-split Linear layers with their rotated inner index -- against the plain convolutions / matrix products (integers, exact)."""
    rng = np.random.default_rng(5)
    W3P, W4P, PTP, A3P = 516, 592, 36, 20
    act2 = rng.integers(-3, 4, size=(32, 12, 12)).astype(np.int64)
    w3 = rng.integers(-2, 3, size=(64, 32, 4, 4)).astype(np.int64)
    b3 = rng.integers(-5, 6, size=64).astype(np.int64)
    w4 = rng.integers(-2, 3, size=(128, 64, 3, 3)).astype(np.int64)
    b4 = rng.integers(-5, 6, size=128).astype(np.int64)
    f0 = rng.integers(-2, 3, size=(64, 128)).astype(np.int64); g0 = rng.integers(-3, 4, size=64).astype(np.int64)
    f2 = rng.integers(-2, 3, size=(32, 64)).astype(np.int64); g2 = rng.integers(-3, 4, size=32).astype(np.int64)
    NA = 9
    f4 = rng.integers(-2, 3, size=(NA, 32)).astype(np.int64); g4 = rng.integers(-3, 4, size=NA).astype(np.int64)

    # ---- reference: conv (valid) + ReLU + floor-mode 2x2 max pool, twice; three Linear layers
    def conv_relu_pool(x, w, b):
        co, ci, k, _ = w.shape
        ho = x.shape[1] - k + 1
        y = np.zeros((co, ho, ho), np.int64)
        for ky in range(k):
            for kx in range(k):
                y += np.einsum("oc,chw->ohw", w[:, :, ky, kx], x[:, ky:ky + ho, kx:kx + ho])
        y = np.maximum(y + b[:, None, None], 0)
        hp = ho // 2
        return y[:, :2 * hp, :2 * hp].reshape(co, hp, 2, hp, 2).max(axis=(2, 4))
    ref3 = conv_relu_pool(act2, w3, b3)                       # (64,4,4)
    ref4 = conv_relu_pool(ref3, w4, b4).reshape(128)          # (128,)
    h1 = np.maximum(f0 @ ref4 + g0, 0); h2 = np.maximum(f2 @ h1 + g2, 0); ref_z = f4 @ h2 + g4

    a2 = act2.reshape(-1)
    tid = np.arange(256)
    lane, warp = tid & 31, tid >> 5
    s_a3 = np.full((8, 64 * A3P), -10 ** 9, np.int64)         # one copy per CTA of the cluster
    for r in range(8):
        # weight slice of CTA r with the padded channel pitch (float4 index i = co * 128 + rest -> co * (W3P / 4) + rest)
        src = w3[8 * r:8 * r + 8].reshape(-1)
        sw3 = np.full(8 * W3P, 10 ** 9, np.int64)
        i4 = np.arange(1024)
        dst = (i4 >> 7) * (W3P // 4) + (i4 & 127)
        for q in range(4):
            sw3[4 * dst + q] = src[4 * i4 + q]
        cp, py, cs = tid & 3, (tid >> 2) & 3, 2 * warp + (lane >> 4)
        assert np.array_equal(cs, tid >> 4)
        acc = np.zeros((256, 2, 2, 8), np.int64)
        for i in range(2):
            ci = 2 * cs + i
            base = ci * 144 + (2 * py) * 12
            inn = a2[base[:, None] + np.arange(60)[None, :]].reshape(256, 5, 12)
            for h in range(2):
                wb = (2 * cp + h) * W3P + ci * 16
                w = sw3[wb[:, None] + np.arange(16)[None, :]].reshape(256, 4, 4)
                for ky in range(4):
                    for dy in range(2):
                        for ox in range(8):
                            acc[:, h, dy, ox] += (w[:, ky, :] * inn[:, dy + ky, ox:ox + 4]).sum(1)
        part = np.full(16 * 16 * PTP, 10 ** 9, np.int64)
        pb = (cs * 16 + (tid & 15)) * PTP
        part[pb[:, None] + np.arange(32)[None, :]] = acc.reshape(256, 32)
        # reduction: thread = (px, dy, h, g); the dy partner is lane ^ 4, the row of four px is lanes ^1 ^2 ^3
        px, dy, h, g = tid & 3, (tid >> 2) & 1, (tid >> 3) & 1, tid >> 4
        cl, ppy = 2 * (g & 3) + h, g >> 2
        pp = g * PTP + h * 16 + dy * 8 + 2 * px
        s0 = sum(part[pp + k * 16 * PTP] for k in range(16)); s1 = sum(part[pp + k * 16 * PTP + 1] for k in range(16))
        v = np.maximum(s0, s1)
        v = np.maximum(v, v[tid ^ 4])
        v = np.maximum(v + b3[8 * r + cl], 0)
        row = np.stack([v, v[tid ^ 1], v[tid ^ 2], v[tid ^ 3]], 1)
        sel = (px == 0) & (dy == 0)
        off = (8 * r + cl) * A3P + ppy * 4
        s_a3[r][off[sel, None] + np.arange(4)[None, :]] = row[sel]
    own = s_a3.copy()
    for r in range(8):                                        # pull: thread -> (source CTA d, channel cl, pooled row py), 16 B each
        d, cl, py = tid >> 5, (tid >> 2) & 7, tid & 3
        off = (8 * d + cl) * A3P + py * 4
        for t in range(256):
            if d[t] != r:
                s_a3[r][off[t]:off[t] + 4] = own[d[t]][off[t]:off[t] + 4]
    for r in range(8):
        assert np.array_equal(s_a3[r].reshape(64, A3P)[:, :16].reshape(64, 4, 4), ref3), r

    s_a4 = np.zeros(128, np.int64)
    for r in range(8):
        src = w4[16 * r:16 * r + 16].reshape(-1)
        sw4 = np.full(16 * W4P, 10 ** 9, np.int64)
        i4 = np.arange(2304)
        dst = (i4 // 144) * (W4P // 4) + (i4 % 144)
        for q in range(4):
            sw4[4 * dst + q] = src[4 * i4 + q]
        s, c = tid & 15, tid >> 4
        o = np.zeros((256, 4), np.int64)
        for i in range(4):
            ci = s + 16 * i
            inn = s_a3[r][(ci * A3P)[:, None] + np.arange(16)[None, :]].reshape(256, 4, 4)
            wp = c * W4P + ci * 9
            for ky in range(3):
                for kx in range(3):
                    w = sw4[wp + ky * 3 + kx]
                    o[:, 0] += w * inn[:, ky, kx]; o[:, 1] += w * inn[:, ky, kx + 1]
                    o[:, 2] += w * inn[:, ky + 1, kx]; o[:, 3] += w * inn[:, ky + 1, kx + 1]
        for offx in (8, 4, 2, 1):
            o = o + o[tid ^ offx]
        feat = np.maximum(o.max(1) + b4[16 * r + c], 0)
        s_a4[16 * r + c[s == 0]] = feat[s == 0]
    assert np.array_equal(s_a4, ref4)

    # head on CTA 0: rows split over neighbouring lanes, inner index rotated per lane
    j, q = tid >> 2, tid & 3
    p = np.zeros(256, np.int64)
    for i in range(8):
        k = (i + lane) & 7
        for e in range(4):
            p += f0.reshape(-1)[j * 128 + q * 32 + 4 * k + e] * s_a4[q * 32 + 4 * k + e]
    p = p + p[tid ^ 1]; p = p + p[tid ^ 2]
    s_h1 = np.zeros(64, np.int64); s_h1[j[q == 0]] = np.maximum(p + g0[j], 0)[q == 0]
    assert np.array_equal(s_h1, h1)
    j, e8 = tid >> 3, tid & 7
    k0 = (e8 >> 2) & 1
    p = np.zeros(256, np.int64)
    for kk in (k0, k0 ^ 1):
        for e in range(4):
            p += f2.reshape(-1)[j * 64 + e8 * 8 + 4 * kk + e] * s_h1[e8 * 8 + 4 * kk + e]
    for offx in (1, 2, 4):
        p = p + p[tid ^ offx]
    s_h2 = np.zeros(32, np.int64); s_h2[j[e8 == 0]] = np.maximum(p + g2[j], 0)[e8 == 0]
    assert np.array_equal(s_h2, h2)
    c, e8 = tid >> 3, tid & 7
    cc = np.minimum(c, NA - 1)
    p = np.zeros(256, np.int64)
    for i in range(4):
        k = (i + c) & 3
        p += f4.reshape(-1)[cc * 32 + e8 * 4 + k] * s_h2[e8 * 4 + k]
    for offx in (1, 2, 4):
        p = p + p[tid ^ offx]
    z = np.zeros(NA, np.int64)
    ok = (e8 == 0) & (c < NA)
    z[c[ok]] = (p + g4[cc])[ok]
    assert np.array_equal(z, ref_z)
    # first maximum: ballot of the lanes that hold the maximum, lowest set bit
    m = z.max()
    assert int(np.flatnonzero(z == m)[0]) == int(np.argmax(z))
