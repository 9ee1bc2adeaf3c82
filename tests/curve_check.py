"""What "the 1k-step loss curves agree" can mean, and the check both arithmetic modes go through.

Adam at lr 1e-3 (imitation.py:83) leaves the ln 9 plateau through a saddle: rounding-sized differences
grow ~10x per 20 steps there (measured below), so two *correct* f32 runs are different trajectories after
step ~100. tests/golden/ref_curve_b8_1k_ensemble.npz (oracle/make_golden.py --ensemble) holds the
UNMODIFIED reference re-run 8 times with every initial weight moved by ~1 f32 ulp and 1/2/4/8 intra-op
threads: those members differ from the canonical reference run by 1e-5 at step 15, 5e-4 at step 35,
1e-2 at step 95, leave the plateau anywhere between steps 95 and 114, have whole-curve means between
1.2777 and 1.3079 (2.4 %) and fall up to 2.5 % outside each other's 50-step-window envelope.
The device curve is therefore held to:
  (1) the same trajectory per step while trajectories are still comparable (caller: first 70 / 30 steps);
  (2) whole-curve mean inside the ensemble's range of means widened by `tol`;
  (3) every 50-step window from step 300 on (all members are past the plateau) inside the ensemble envelope
      widened by `tol` + the ensemble's own leave-one-out excess;
  (4) plateau exit (first step whose 20-step running mean is < 2.0) within `exit_slack` steps of the ensemble's;
  (5) the last 250 steps' mean inside the ensemble's range widened by `tol`.
`tol` is the north star's 1 % (fp32 mode) / 2 % (bf16 mode).

bf16 mode: the saddle escape amplifies bf16-sized rounding (4e-3 relative on the gradients) into +-100 steps of exit
time, or no exit within 1000 steps -- measured on the UNMODIFIED reference with that noise injected into its f32
gradients (2 of 6 runs never left the plateau). So the bf16 curve is checked (a) per step on the
plateau from the common initial state and (b) for steps 200..999 from the reference's own parameters and Adam state
at step 200, where the dynamics are contractive again and "the curves agree within 2 %" is a meaningful statement."""
import os

import numpy as np


def _exit_step(curve):
    m = np.convolve(curve, np.ones(20) / 20, "valid")
    return int(np.argmax(m < 2.0))


def check_curve(got, golden_dir, tol, exit_slack=25, start=0):
    """`got` = losses of steps [start, 1000). start > 0 is a run resumed from the reference's own state at that step
    (tests/golden/ref_state_b8_step200.npz): checks (2) and (4), which are about the plateau, do not apply to it."""
    ref = np.load(os.path.join(golden_dir, "ref_curve_b8_1k.npz"))["losses"]
    ens = np.load(os.path.join(golden_dir, "ref_curve_b8_1k_ensemble.npz"))["losses"]
    members = np.vstack([ens, ref[None]])                      # 9 runs of the unmodified reference
    assert start % 50 == 0 and len(got) == 1000 - start
    got = np.concatenate([np.full(start, np.nan), got])
    w = members.reshape(len(members), -1, 50).mean(2)
    excess = 0.0                                               # how far a member falls outside the OTHER members' envelope
    for m in range(len(w)):
        oth = np.delete(w, m, 0)
        excess = max(excess, float((np.maximum(oth.min(0) - w[m], w[m] - oth.max(0)).clip(0) / w[m])[6:].max()))
    gw = got.reshape(-1, 50).mean(1)
    out = np.maximum(w.min(0) - gw, gw - w.max(0)).clip(0) / gw
    assert out[6:].max() <= tol + excess, (float(out[6:].max()), excess, int(out[6:].argmax()) + 6)
    for lo_step in (750, 300):                                 # (5) and the whole post-plateau part
        stat = np.sort(members[:, lo_step:].mean(1))
        loo = max(stat[1] / stat[0], stat[-1] / stat[-2]) - 1.0   # leave-one-out excess of this statistic
        g = got[lo_step:].mean()
        assert stat[0] * (1 - tol - loo) <= g <= stat[-1] * (1 + tol + loo), (lo_step, g, stat, loo)
    if start == 0:
        means = members.mean(1)
        assert means.min() - tol * ref.mean() <= got.mean() <= means.max() + tol * ref.mean(), (got.mean(), means)
        exits = [_exit_step(c) for c in members]
        assert min(exits) - exit_slack <= _exit_step(got) <= max(exits) + exit_slack, (_exit_step(got), exits)
